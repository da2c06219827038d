"""Drop-in replacements for the three free functions of the reference's ops layer
(/root/reference/src/utils/ops.py:183-202), same names, same argument meaning:

    resample(image, flow, mode='bilinear')      ops.py:187
    grid_sample(input1, input2, mode='bilinear') ops.py:183
    get_grid(batchsize, rows, cols, gpu_id=0)    ops.py:196

`resample` no longer builds a grid on the CPU, copies it to the GPU and runs five kernels: it is
one launch of the fused sm_100a kernel (c2m_b200.functional.warp_blend with mask=None).
"""
from __future__ import annotations

import torch

from . import _lib
from .functional import WarpBlendFunction, deterministic_default, warp_blend

__all__ = ["resample", "grid_sample", "get_grid", "warp_blend", "get_occlusion_map", "get_corresponding_map"]


def _check_mode(mode: str) -> None:
    # every call site of the reference passes the default; ATen's other modes
    # ('nearest', 'bicubic') are not part of this path
    if mode != "bilinear":
        raise NotImplementedError(f"c2m_b200: only mode='bilinear' is implemented, got {mode!r}")


def resample(image: torch.Tensor, flow: torch.Tensor, mode: str = "bilinear") -> torch.Tensor:
    """Backward-warp `image` [B,C,H,W] by the pixel `flow` [B,2,H,W] (channel 0 = x) with the
    reference's exact convention: (size-1)/2 flow normalisation, align_corners=False sampling,
    border padding (ops.py:187-193)."""
    _check_mode(mode)
    return warp_blend(image, flow, None)


def grid_sample(input1: torch.Tensor, input2: torch.Tensor, mode: str = "bilinear") -> torch.Tensor:
    """F.grid_sample(input1, input2, padding_mode='border') of ops.py:183-184: `input2` is a
    normalised sampling grid [N,H,W,2]."""
    _check_mode(mode)
    if input2.dim() != 4 or input2.shape[-1] != 2:
        raise ValueError(f"grid must be [N,H,W,2], got {tuple(input2.shape)}")
    if tuple(input2.shape[1:3]) != tuple(input1.shape[2:]):
        # ATen allows an output size different from the input size; no call site of the path does
        raise NotImplementedError("c2m_b200.grid_sample: grid and input must share H and W")
    return _GridSampleFunction.apply(input1, input2)


class _GridSampleFunction(torch.autograd.Function):
    # under autocast the op runs in float32 with its inputs cast up, like WarpBlendFunction
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, image, grid):
        if not (image.is_cuda and grid.is_cuda):
            raise RuntimeError("c2m_b200.grid_sample: CUDA tensors required (no CPU fallback)")
        if image.dtype != torch.float32 or grid.dtype != torch.float32:
            raise TypeError(f"c2m_b200.grid_sample: float32 tensors required, got {image.dtype} and {grid.dtype}")
        if image.device != grid.device:
            raise RuntimeError("c2m_b200.grid_sample: all tensors must be on the same device")
        if image.dim() != 4:
            raise ValueError(f"expected an image [B,C,H,W], got {tuple(image.shape)}")
        if image.shape[0] != grid.shape[0] and (image.shape[0] == 0 or grid.shape[0] % image.shape[0] != 0):
            raise ValueError(f"image batch {image.shape[0]} must equal or divide the grid batch {grid.shape[0]}")
        x = image.contiguous()
        g = grid.contiguous()  # [N,H,W,2]; the kernels take it through the `flow` slot
        N, H, W, _ = g.shape
        C = x.shape[1]
        out = torch.empty((N, C, H, W), dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            _lib.warp_blend_fwd(x.data_ptr(), g.data_ptr(), None, None, out.data_ptr(), N, C, H, W, x.shape[0],
                                x.stride(), out.stride(), _lib.PAD_BORDER, _lib.FLAG_COORD_GRID,
                                torch.cuda.current_stream().cuda_stream)
        ctx.save_for_backward(x, g)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gout):
        x, g = ctx.saved_tensors
        gout = gout.contiguous()
        N, H, W, _ = g.shape
        C = x.shape[1]
        need_x, need_g = ctx.needs_input_grad
        gx = torch.empty_like(x) if need_x else None
        gg = torch.empty_like(g) if need_g else None
        flags = _lib.FLAG_COORD_GRID | (_lib.FLAG_DETERMINISTIC if deterministic_default() else 0)
        with torch.cuda.device(x.device):
            nbytes = _lib.bwd_workspace_bytes(N, C, H, W, x.shape[0], need_x, flags)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
            _lib.warp_blend_bwd(x.data_ptr(), g.data_ptr(), None, None, gout.data_ptr(),
                                None if gx is None else gx.data_ptr(), None if gg is None else gg.data_ptr(),
                                None, None, N, C, H, W, x.shape[0], x.stride(), gout.stride(), _lib.PAD_BORDER,
                                flags, ws.data_ptr(), nbytes, torch.cuda.current_stream().cuda_stream)
        return gx, gg


def get_grid(batchsize: int, rows: int, cols: int, gpu_id=0) -> torch.Tensor:
    """[B,2,rows,cols] base grid of ops.py:196-202, produced on the device (no CPU build, no H2D
    copy) and bit-identical to the reference's CPU float32 linspace construction."""
    device = gpu_id if isinstance(gpu_id, torch.device) else torch.device("cuda", int(gpu_id))
    grid = torch.empty((batchsize, 2, rows, cols), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        _lib.base_grid(grid.data_ptr(), batchsize, rows, cols, torch.cuda.current_stream().cuda_stream)
    return grid


def _splat(data: torch.Tensor, flags: int) -> torch.Tensor:
    if data.dim() != 4 or data.shape[1] != 2:
        raise ValueError(f"expected [B,2,H,W], got {tuple(data.shape)}")
    if not data.is_cuda or data.dtype != torch.float32:
        raise RuntimeError("c2m_b200: float32 CUDA tensors required (no CPU fallback)")
    d = data.detach().contiguous()
    B, _, H, W = d.shape
    out = torch.empty((B, 1, H, W), dtype=torch.float32, device=d.device)
    with torch.cuda.device(d.device):
        nbytes = _lib.occlusion_map_workspace_bytes(B, H, W)
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=d.device)
        _lib.occlusion_map(d.data_ptr(), out.data_ptr(), B, H, W, flags, ws.data_ptr(), nbytes,
                           torch.cuda.current_stream().cuda_stream)
    return out


def get_corresponding_map(data: torch.Tensor) -> torch.Tensor:
    """ops.py:205-251: `data` [B,2,H,W] absolute pixel coordinates -> [B,1,H,W] sum of the bilinear splat weights
    that land on each pixel (corners outside the image dropped).  One scatter kernel with order-independent
    fixed-point accumulation instead of eight temporaries + scatter_add_."""
    return _splat(data, _lib.OCC_COORDS | _lib.OCC_NO_CLAMP)


def get_occlusion_map(flow: torch.Tensor) -> torch.Tensor:
    """ops.py:263-275: forward-splat the pixel `flow` [B,2,H,W] and clamp to [0, 1] (0 = occluded / nothing lands
    there).  Like the reference it carries no gradient (ops.py:271 runs under no_grad)."""
    return _splat(flow, 0)


# re-exported so `from c2m_b200.ops import *` mirrors `from utils.ops import *`
_ = WarpBlendFunction
