"""torch.autograd.Function over the C ABI: the fused flow-warp + occlusion-blend op.

Contract in reference terms (PierfrancescoArdino/C2M):
    warp_blend(x, flow, mask) == utils.resample(x, flow) * mask
        src/utils/ops.py:187-193, src/modules/generator/generator.py:93
on CUDA float32 tensors, forward and backward, in one kernel each way.  PyTorch is used for
device memory, streams and autograd bookkeeping only; the arithmetic runs in libc2m_warp.so.
"""
from __future__ import annotations

import os

import torch
from torch.autograd.function import once_differentiable

from . import _lib

_PADDING = {"border": _lib.PAD_BORDER, "zeros": _lib.PAD_ZEROS}


def deterministic_default() -> bool:
    """Deterministic grad-input is selected by torch's global switch or C2M_WARP_DETERMINISTIC=1
    (no new YAML keys in the reference's config)."""
    if os.environ.get("C2M_WARP_DETERMINISTIC", "0") not in ("", "0"):
        return True
    return torch.are_deterministic_algorithms_enabled()


def _ptr(t):
    return None if t is None else t.data_ptr()


try:  # the raw handle of the current stream without building a torch.cuda.Stream object (4 us per call on the host,
    _raw_stream = torch._C._cuda_getCurrentRawStream  # which the small pyramid levels feel)
except AttributeError:  # pragma: no cover
    _raw_stream = None


def _current_stream_ptr(device: torch.device) -> int:
    if _raw_stream is not None:
        return _raw_stream(device.index if device.index is not None else torch.cuda.current_device())
    return torch.cuda.current_stream(device).cuda_stream


class _on_device:
    """`with torch.cuda.device(dev)` only when `dev` is not already current (the switch costs ~10 us of host time
    per call, which is most of what the small pyramid levels spend)."""

    def __init__(self, dev):
        self.ctx = None if dev.index is None or dev.index == torch.cuda.current_device() else torch.cuda.device(dev)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


def _is_nhwc_dense(t: torch.Tensor) -> bool:
    return (t.dim() == 4 and not t.is_contiguous() and t.is_contiguous(memory_format=torch.channels_last))


def _dense(t: torch.Tensor, nhwc: bool) -> torch.Tensor:
    return t.contiguous(memory_format=torch.channels_last) if nhwc else t.contiguous()


def _nchw_policy() -> str:
    """What happens to an NCHW-contiguous `x` (the layout the reference's convolutions produce):
    "propagate" (default) -- x is converted to channels-last ONCE in the forward (c2m_relayout), that copy is what the
        backward keeps, and out / grad-input come back channels-last strided (same shapes and values; the memory
        format then propagates through the caller's convolutions like any channels_last tensor);
    "strict" (C2M_WARP_NCHW=strict or flags=FLAG_STRICT_LAYOUT) -- results keep x's strides: NCHW forward kernel,
        backward staged through three channels-last copies in the workspace."""
    return os.environ.get("C2M_WARP_NCHW", "propagate")


_NO_PROMOTE = (_lib.FLAG_FORCE_GENERIC | _lib.FLAG_COORD_GRID | _lib.FLAG_ALIGN_CORNERS | _lib.FLAG_BWD_ATOMIC | _lib.FLAG_NO_STAGE |
               _lib.FLAG_STRICT_LAYOUT | _lib.FLAG_TRUE_DIV | _lib.FLAG_NO_FMA)


def _relayout_ok(x: torch.Tensor) -> bool:
    C = x.shape[1]
    return x.numel() > 0 and x.shape[0] <= 65535 and C >= 8 and C % 4 == 0 and 4 * x.numel() <= _STAGE_MAX_BYTES


def _promotes(x: torch.Tensor, flags: int) -> bool:
    return _relayout_ok(x) and not (flags & _NO_PROMOTE) and _nchw_policy() != "strict"


def _to_channels_last(x: torch.Tensor) -> torch.Tensor:
    """NCHW-contiguous -> channels-last copy through the library's own relayout kernel (128-bit both ways)."""
    B, C, H, W = x.shape
    y = torch.empty_like(x, memory_format=torch.channels_last)
    with _on_device(x.device):
        _lib.relayout(x.data_ptr(), y.data_ptr(), B, C, H, W, True, _current_stream_ptr(x.device))
    return y


# Plan (opt-in, C2M_WARP_PLAN=1): the backward's segment registration depends on the flow and the mask only, so the
# forward call can run it on a second stream (include/c2m_warp.h, c2m_warp_plan) and the backward starts with its
# gather kernel.  Next to this library's own forward kernel the overlap is small (the registration's SM time does not
# vanish: -10 us of a 1.50 ms step) and the plan holds 24 B per output pixel until the backward,
# hence off by default; a caller that can run c2m_warp_plan under unrelated work gains the whole 0.06 ms.  Levels below
# C2M_WARP_PLAN_MIN_PIXELS output pixels are host-bound and never planned.
_PLAN_MIN_PIXELS = int(os.environ.get("C2M_WARP_PLAN_MIN_PIXELS", str(1 << 20)))
_side_streams = {}


def _plan_enabled() -> bool:
    return os.environ.get("C2M_WARP_PLAN", "0") not in ("", "0")


def _side_stream(device: torch.device) -> "torch.cuda.Stream":
    s = _side_streams.get(device.index)
    if s is None:
        s = _side_streams[device.index] = torch.cuda.Stream(device=device)
    return s


_RESIZE_MODES = {"half_pixel": _lib.RESIZE_HALF_PIXEL, "corners_rescaled": _lib.RESIZE_CORNERS_RESCALE}

# NCHW tensors are staged through channels-last copies in the backward workspace (three tensor-sized buffers) only
# while those copies stay below this many bytes; larger calls run the NCHW kernels (slower, no extra memory).
_STAGE_MAX_BYTES = int(os.environ.get("C2M_WARP_STAGE_MAX_BYTES", str(8 << 30)))
# Backward workspaces up to this size are kept per (device, stream) and reused by later calls on that stream (stream
# order makes that safe); larger ones come from torch's caching allocator per call.
_WS_CACHE_MAX_BYTES = int(os.environ.get("C2M_WARP_WS_CACHE_BYTES", str(64 << 20)))
_ws_cache = {}


def _workspace(nbytes: int, device: torch.device, stream) -> torch.Tensor:
    if nbytes > _WS_CACHE_MAX_BYTES or torch.cuda.is_current_stream_capturing():
        # (a buffer captured into a CUDA graph must not be recycled by later eager calls)
        return torch.empty(nbytes, dtype=torch.uint8, device=device)
    key = (device.index, stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def _resize_spec(x, flow, mask, flow_resize):
    """c2m_resize for a flow / mask held at another resolution than x, or as 5-D clips (None: plain [N,c,H,W])."""
    H, W = x.shape[2:]
    fh, fw = flow.shape[-2:]
    mh, mw = (mask.shape[-2:] if mask is not None else (H, W))
    if flow.dim() == 5:  # [B,2,T,h,w] (+ mask [B,1,T,h',w']): frame n = t * B + b, no folded copies
        return _lib.resize_spec(fh, fw, mh, mw, _RESIZE_MODES[flow_resize or "half_pixel"], flow.shape[2])
    if (fh, fw) == (H, W) and (mh, mw) == (H, W):
        return None
    if (fh, fw) != (H, W) and flow_resize is None:
        raise ValueError(f"flow {tuple(flow.shape)} does not match x {tuple(x.shape)} spatially "
                         "(pass flow_resize='half_pixel' or 'corners_rescaled' to resize it inside the kernel)")
    mode = _RESIZE_MODES[flow_resize or "half_pixel"]
    return _lib.resize_spec(fh, fw, mh, mw, mode)


def _check_inputs(x, flow, mask, other, resized=False):
    for name, t in (("x", x), ("flow", flow), ("mask", mask), ("other", other)):
        if t is None:
            continue
        if not t.is_cuda:
            # the reference itself cannot run this path on CPU tensors (ops.py:189,202)
            raise RuntimeError(f"c2m_b200.warp_blend: `{name}` must be a CUDA tensor (no CPU fallback)")
        if t.dtype != torch.float32:
            raise TypeError(f"c2m_b200.warp_blend: `{name}` must be float32, got {t.dtype}")
        if t.device != x.device:
            raise RuntimeError("c2m_b200.warp_blend: all tensors must be on the same device")
    if x.dim() != 4 or flow.dim() not in (4, 5) or flow.shape[1] != 2:
        raise ValueError(f"expected x [B,C,H,W] and flow [N,2,H,W], got {tuple(x.shape)} and {tuple(flow.shape)}")
    if flow.dim() == 5:  # 5-D clips [B',2,T,h,w] / [B',1,T,h',w']
        N = flow.shape[0] * flow.shape[2]
        if mask is not None and (mask.dim() != 5 or mask.shape[:3] != (flow.shape[0], 1, flow.shape[2])):
            raise ValueError(f"a 5-D flow {tuple(flow.shape)} needs a 5-D mask [B,1,T,h,w], got {tuple(mask.shape)}")
        if other is not None:
            raise ValueError("`other` is not supported together with 5-D clips")
    else:
        N = flow.shape[0]
        if mask is not None and (mask.dim() != 4 or mask.shape[0] != N or mask.shape[1] != 1):
            raise ValueError(f"mask must be [N,1,h,w] with N={N}, got {tuple(mask.shape)}")
    H, W = x.shape[2:]
    B = x.shape[0]
    if B != N and (B == 0 or N % B != 0):
        raise ValueError(f"x batch {B} must equal or divide the flow batch {N}")
    if not resized:
        if tuple(flow.shape[2:]) != (H, W):
            raise ValueError(f"flow {tuple(flow.shape)} does not match x {tuple(x.shape)} spatially")
        if mask is not None and tuple(mask.shape) != (N, 1, H, W):
            raise ValueError(f"mask must be [N,1,H,W]={N, 1, H, W}, got {tuple(mask.shape)}")
    elif min(flow.shape[-2:]) == 0 or (mask is not None and min(mask.shape[-2:]) == 0):
        raise ValueError("cannot resize an empty flow / mask")
    if other is not None:
        if mask is None:
            raise ValueError("`other` requires a mask")
        if tuple(other.shape) != (N, x.shape[1], H, W):
            raise ValueError("`other` must have the output's shape")


class WarpBlendFunction(torch.autograd.Function):
    """out = mask * bilinear_border_warp(x, flow) [+ (1 - mask) * other]."""

    # under autocast the op runs in float32 with its inputs cast up, the convention of the reference's own
    # native warp (src/modules/third_party/resample2d/resample2d.py:55-57: autocast(False) + .float())
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, flow, mask, other, padding, deterministic, flags, flow_resize=None):
        if x.dim() != 4 or flow.dim() not in (4, 5):
            raise ValueError(f"expected x [B,C,H,W] and flow [N,2,h,w], got {tuple(x.shape)} and {tuple(flow.shape)}")
        rs = _resize_spec(x, flow, mask, flow_resize)
        _check_inputs(x, flow, mask, other, resized=rs is not None)
        nhwc = _is_nhwc_dense(x)
        x = _dense(x, nhwc)
        promote = not nhwc and _promotes(x, flags)
        flags &= ~_lib.FLAG_STRICT_LAYOUT
        flow = flow.contiguous()
        mask = None if mask is None else mask.contiguous()
        N = flow.shape[0] * (flow.shape[2] if flow.dim() == 5 else 1)
        H, W = x.shape[2:]
        B, C = x.shape[0], x.shape[1]
        plan = side = None
        with _on_device(x.device):
            cur = None
            stream = _current_stream_ptr(x.device)
            if ((nhwc or promote) and rs is None and not deterministic and ctx.needs_input_grad[0] and _plan_enabled()
                    and N * H * W >= _PLAN_MIN_PIXELS):
                nbytes = _lib.plan_bytes(N, C, H, W, B, flags)
                if nbytes:
                    plan = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
                    cur = torch.cuda.current_stream()
                    side = _side_stream(x.device)
                    side.wait_stream(cur)  # the flow / mask (and the buffer's previous life) belong to `cur`
            plan_first = promote
            if promote:
                # NCHW x: the channels-last copy is what the backward keeps.  The plan (integer-bound) runs next to
                # the conversion (HBM-bound, few registers) and is done before the forward kernel starts
                x, nhwc = _to_channels_last(x), True
            if plan is not None and plan_first:
                _lib.warp_plan(_ptr(flow), _ptr(mask), N, C, H, W, B, padding, flags, _ptr(plan), nbytes,
                               side.cuda_stream)
            other = None if other is None else _dense(other, nhwc)
            out = torch.empty((N, C, H, W), dtype=x.dtype, device=x.device,
                              memory_format=torch.channels_last if nhwc else torch.contiguous_format)
            _lib.warp_blend_fwd(_ptr(x), _ptr(flow), _ptr(mask), _ptr(other), _ptr(out), N, C, H, W, B,
                                x.stride(), out.stride(), padding, flags, stream, rs)
            if plan is not None:
                if not plan_first:
                    # launched after the forward kernel, which keeps the scheduler's priority: the plan's small blocks
                    # take the registers and thread slots that kernel leaves free on every SM
                    _lib.warp_plan(_ptr(flow), _ptr(mask), N, C, H, W, B, padding, flags, _ptr(plan), nbytes,
                                   side.cuda_stream)
                cur.wait_stream(side)  # everything later on `cur` -- the backward, the buffer's release -- is ordered
        ctx.save_for_backward(x, flow, mask, other)  # inputs only: geometry is recomputed in backward
        ctx.cfg = (padding, bool(deterministic), flags, nhwc, rs)
        ctx.plan = plan
        return out

    @staticmethod
    @once_differentiable
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gout):
        x, flow, mask, other = ctx.saved_tensors
        padding, deterministic, flags, nhwc, rs = ctx.cfg
        need_x, need_flow, need_mask, need_other = ctx.needs_input_grad[:4]
        need_mask = need_mask and mask is not None
        need_other = need_other and other is not None
        if nhwc and gout.is_contiguous() and not _is_nhwc_dense(gout) and _relayout_ok(gout):
            # an NCHW upstream gradient for a channels-last result: the library's relayout kernel (copy speed; torch's
            # strided copy takes three times as long)
            gout = _to_channels_last(gout)
        else:
            gout = _dense(gout, nhwc)
        N = flow.shape[0] * (flow.shape[2] if flow.dim() == 5 else 1)
        H, W = x.shape[2:]
        B, C = x.shape[0], x.shape[1]
        gx = torch.empty_like(x) if need_x else None
        gflow = torch.empty_like(flow) if need_flow else None
        gmask = torch.empty_like(mask) if need_mask else None
        gother = torch.empty_like(gout) if need_other else None
        if deterministic:
            flags |= _lib.FLAG_DETERMINISTIC
        if (not nhwc and need_x and C % 4 == 0 and not (flags & _lib.FLAG_NO_STAGE)
                and 4 * (N + 2 * B) * C * H * W <= _STAGE_MAX_BYTES):
            # NCHW tensors: the library stages them through channels-last copies in the workspace and runs
            # the channels-last kernels (about twice as fast as gathering 4-byte elements at NCHW strides);
            # C2M_WARP_STAGE_MAX_BYTES bounds the extra footprint
            flags |= _lib.FLAG_STAGE_NHWC
        plan, ctx.plan = ctx.plan, None  # a plan serves one backward (a second one over a retained graph bins again)
        with _on_device(x.device):
            stream = _current_stream_ptr(x.device)
            if (plan is not None and need_x and not (flags & _lib.FLAG_DETERMINISTIC)
                    and (gout.data_ptr() | gx.data_ptr() | x.data_ptr()) % 16 == 0):
                flags |= _lib.FLAG_PLANNED
                ws, ws_bytes = plan, plan.numel()
            else:
                ws_bytes = _lib.bwd_workspace_bytes(N, C, H, W, B, need_x, flags, rs)
                ws = _workspace(ws_bytes, x.device, stream)
            _lib.warp_blend_bwd(_ptr(x), _ptr(flow), _ptr(mask), _ptr(other), _ptr(gout), _ptr(gx), _ptr(gflow),
                                _ptr(gmask), _ptr(gother), N, C, H, W, B, x.stride(), gout.stride(), padding,
                                flags, _ptr(ws), ws_bytes, stream, rs)
        return gx, gflow, gmask, gother, None, None, None, None


def warp_blend(x, flow, mask=None, other=None, padding="border", deterministic=None, flags=0, flow_resize=None,
               align_corners=False):
    """Fused ``resample(x, flow) * mask`` of the reference (ops.py:187-193, generator.py:93).

    x      [B,C,H,W] float32 CUDA, channels-last or NCHW-contiguous.  The results follow a channels-last x; an NCHW
           x with C >= 8, C % 4 == 0 is converted once in the forward and out / grad-input come back channels-last
           strided (same shape and values; C2M_WARP_NCHW=strict or flags=FLAG_STRICT_LAYOUT keeps x's strides);
           B == N, or B divides N and frame n samples image n % B (the T-fold repeat of
           motion_autoencoder.py:117-119 without materialising it).
    flow   [N,2,H,W] displacement in pixels, channel 0 = x.
    mask   [N,1,H,W] or None (None: plain warp, generator.py:95-96).
    other  optional [N,C,H,W]: out = mask*warp + (1-mask)*other (not in the reference).
    flow_resize  flow [N,2,h,w] / mask [N,1,h',w'] held at another resolution than x are resized to (H, W) inside
           the kernel (bilinear; the gradients come back at their own sizes): "half_pixel" = generator.py:84-85
           (align_corners=False, values kept), "corners_rescaled" = utils.py:346-354 (align_corners=True, values
           scaled by new/old).  The mask is always resized with align_corners=False (generator.py:92).
           flow / mask may also be the reference's 5-D clips [B',2,T,h,w] / [B',1,T,h',w']: frame n = t * B' + b reads
           plane (b, :, t) -- the fold torch.cat(torch.unbind(., 2), 0) of motion_autoencoder.py:120-123 without the
           copies -- and their gradients come back 5-D.
    align_corners  True samples with F.grid_sample's align_corners=True convention (zero flow is then the identity; no
           call site of the reference uses it -- stride-generic kernels).
    """
    if deterministic is None:
        deterministic = deterministic_default()
    if align_corners:
        flags = int(flags) | _lib.FLAG_ALIGN_CORNERS
    return WarpBlendFunction.apply(x, flow, mask, other, _PADDING[padding], bool(deterministic), int(flags),
                                   flow_resize)
