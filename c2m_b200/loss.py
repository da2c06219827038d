"""Fused warped-frame L1 loss: the `loss_dict["warped"]` term of the reference's training losses
(/root/reference/src/losses/losses.py:219-222 with L1MaskedLoss, losses.py:184-189, mask None):

    warped_frames = cat([unsqueeze(resample(source_frame, dense_motion_bw[:, :, i]), 2) for i in range(T)], 2)
    loss = F.l1_loss(warped_frames, target_frames)

The reference runs T warps of the C=3 source frame (each of them a CPU-built grid, a host-to-device copy and
five kernels), concatenates the T results and reduces.  `warped_l1_loss` is one kernel forward (+ a one-block
finish) and one kernel backward over the 5-D tensors as they are; the warped clip never exists in memory.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch.autograd.function import once_differentiable

from . import _lib
from .functional import warp_blend

__all__ = ["warped_l1_loss", "WarpedL1Function", "flow_consistency_loss", "FlowConsistencyFunction", "FlowConsistLoss"]


class WarpedL1Function(torch.autograd.Function):
    """mean |resample(source, flows[:, :, t]) - targets[:, :, t]|; gradients for flows and targets."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, source, flows, targets):
        source, flows, targets = source.contiguous(), flows.contiguous(), targets.contiguous()
        B, C, H, W = source.shape
        T = flows.shape[2]
        loss = torch.empty((), dtype=torch.float32, device=source.device)
        with torch.cuda.device(source.device):
            ws_bytes = _lib.warped_l1_workspace_bytes()
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=source.device)
            _lib.warped_l1_fwd(source.data_ptr(), flows.data_ptr(), targets.data_ptr(), loss.data_ptr(), B, C, T, H, W,
                               ws.data_ptr(), ws_bytes, torch.cuda.current_stream().cuda_stream)
        ctx.save_for_backward(source, flows, targets)
        return loss

    @staticmethod
    @once_differentiable
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gloss):
        source, flows, targets = ctx.saved_tensors
        if ctx.needs_input_grad[0]:
            raise RuntimeError("c2m_b200.WarpedL1Function has no gradient for the source frame "
                               "(warped_l1_loss composes warp_blend + l1_loss in that case)")
        B, C, H, W = source.shape
        T = flows.shape[2]
        gloss = gloss.to(torch.float32).contiguous()
        gflows = torch.empty_like(flows) if ctx.needs_input_grad[1] else None
        gtargets = torch.empty_like(targets) if ctx.needs_input_grad[2] else None
        with torch.cuda.device(source.device):
            _lib.warped_l1_bwd(source.data_ptr(), flows.data_ptr(), targets.data_ptr(), gloss.data_ptr(),
                               None if gflows is None else gflows.data_ptr(),
                               None if gtargets is None else gtargets.data_ptr(), B, C, T, H, W,
                               torch.cuda.current_stream().cuda_stream)
        return None, gflows, gtargets


def _check(source, flows, targets):
    for name, t in (("source", source), ("flows", flows), ("targets", targets)):
        if not t.is_cuda:
            raise RuntimeError(f"c2m_b200.warped_l1_loss: `{name}` must be a CUDA tensor (no CPU fallback)")
        if t.dtype != torch.float32 and not torch.is_autocast_enabled():
            raise TypeError(f"c2m_b200.warped_l1_loss: `{name}` must be float32, got {t.dtype}")
        if t.device != source.device:
            raise RuntimeError("c2m_b200.warped_l1_loss: all tensors must be on the same device")
    if source.dim() != 4 or flows.dim() != 5 or targets.dim() != 5 or flows.shape[1] != 2:
        raise ValueError(f"expected source [B,C,H,W], flows [B,2,T,H,W], targets [B,C,T,H,W]; got "
                         f"{tuple(source.shape)}, {tuple(flows.shape)}, {tuple(targets.shape)}")
    B, C, H, W = source.shape
    T = flows.shape[2]
    if tuple(flows.shape) != (B, 2, T, H, W) or tuple(targets.shape) != (B, C, T, H, W):
        raise ValueError(f"shape mismatch: source {tuple(source.shape)}, flows {tuple(flows.shape)}, "
                         f"targets {tuple(targets.shape)}")


def warped_l1_loss(source: torch.Tensor, flows: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
    """F.l1_loss of the source frame warped by each of the T flows against the T target frames
    (losses.py:219-222).  source [B,C,H,W], flows [B,2,T,H,W] in pixels (channel 0 = x), targets [B,C,T,H,W]."""
    _check(source, flows, targets)
    if source.requires_grad and torch.is_grad_enabled():
        # not a case of the reference (the frames are data): the T warps as ONE launch of the fused warp kernel
        # (frame n = t * B + b samples image n % B), then torch's reduction
        B, C, H, W = source.shape
        T = flows.shape[2]
        warped = warp_blend(source, flows.permute(2, 0, 1, 3, 4).reshape(T * B, 2, H, W), None)
        return F.l1_loss(warped.view(T, B, C, H, W).permute(1, 2, 0, 3, 4), targets)
    return WarpedL1Function.apply(source, flows, targets)


# ------------------------------------------------------------------------------------------------------------------
# Flow-consistency loss (losses.py:115-141)
def _p(t):
    return None if t is None else t.data_ptr()


class FlowConsistencyFunction(torch.autograd.Function):
    """scale * (mean(mask_bw*|resample(flow, flowback) + flowback|) + mean(mask_fw*|resample(flowback, flow) + flow|))
    on the 5-D tensors as they are; gradients for both flows and both masks."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, flow, flowback, mask_fw, mask_bw, scale):
        flow, flowback = flow.contiguous(), flowback.contiguous()
        mask_fw = None if mask_fw is None else mask_fw.contiguous()
        mask_bw = None if mask_bw is None else mask_bw.contiguous()
        B, _, T, H, W = flow.shape
        loss = torch.empty((), dtype=torch.float32, device=flow.device)
        with torch.cuda.device(flow.device):
            ws_bytes = _lib.flow_consistency_workspace_bytes(B, T, H, W)
            ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=flow.device)
            _lib.flow_consistency_fwd(flow.data_ptr(), flowback.data_ptr(), _p(mask_fw), _p(mask_bw), loss.data_ptr(),
                                      B, T, H, W, float(scale), ws.data_ptr(), ws_bytes,
                                      torch.cuda.current_stream().cuda_stream)
        ctx.save_for_backward(flow, flowback, mask_fw, mask_bw)
        ctx.scale = float(scale)
        return loss

    @staticmethod
    @once_differentiable
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gloss):
        flow, flowback, mask_fw, mask_bw = ctx.saved_tensors
        B, _, T, H, W = flow.shape
        need = ctx.needs_input_grad
        gloss = gloss.to(torch.float32).contiguous()
        gflow = torch.empty_like(flow) if need[0] else None
        gback = torch.empty_like(flowback) if need[1] else None
        gmfw = torch.empty_like(mask_fw) if (need[2] and mask_fw is not None) else None
        gmbw = torch.empty_like(mask_bw) if (need[3] and mask_bw is not None) else None
        with torch.cuda.device(flow.device):
            ws_bytes = _lib.flow_consistency_workspace_bytes(B, T, H, W)
            ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=flow.device)
            _lib.flow_consistency_bwd(flow.data_ptr(), flowback.data_ptr(), _p(mask_fw), _p(mask_bw), gloss.data_ptr(),
                                      _p(gflow), _p(gback), _p(gmfw), _p(gmbw), B, T, H, W, ctx.scale, ws.data_ptr(),
                                      ws_bytes, torch.cuda.current_stream().cuda_stream)
        return gflow, gback, gmfw, gmbw, None


def flow_consistency_loss(flow, flowback, mask_fw=None, mask_bw=None, num_predicted_frames=None):
    """`FlowConsistLoss.forward` (losses.py:131-141).  flow / flowback [B,2,T,H,W] in pixels, masks [B,1,T,H,W] or
    None (the reference tests `mask_bw is not None` and then uses both).  The result is multiplied by
    `num_predicted_frames` (default: T, the value the shipped YAML makes it)."""
    for name, t in (("flow", flow), ("flowback", flowback), ("mask_fw", mask_fw), ("mask_bw", mask_bw)):
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(f"c2m_b200.flow_consistency_loss: `{name}` must be a CUDA tensor (no CPU fallback)")
        if t.dtype != torch.float32 and not torch.is_autocast_enabled():
            raise TypeError(f"c2m_b200.flow_consistency_loss: `{name}` must be float32, got {t.dtype}")
    if flow.dim() != 5 or flow.shape[1] != 2 or flowback.shape != flow.shape:
        raise ValueError(f"expected flow and flowback [B,2,T,H,W], got {tuple(flow.shape)}, {tuple(flowback.shape)}")
    if mask_bw is None:
        mask_fw = None  # losses.py:132: only `mask_bw is not None` selects the masked form
    else:
        B, _, T, H, W = flow.shape
        if mask_fw is None or tuple(mask_fw.shape) != (B, 1, T, H, W) or tuple(mask_bw.shape) != (B, 1, T, H, W):
            raise ValueError("mask_fw and mask_bw must both be [B,1,T,H,W]")
    scale = flow.shape[2] if num_predicted_frames is None else num_predicted_frames
    return FlowConsistencyFunction.apply(flow, flowback, mask_fw, mask_bw, float(scale))


class FlowConsistLoss(torch.nn.Module):
    """Drop-in for the reference module of the same name (losses.py:115-141): same constructor, same forward."""

    def __init__(self, train_params):
        super().__init__()
        self.train_params = train_params

    def forward(self, flow, flowback, mask_fw=None, mask_bw=None):
        return flow_consistency_loss(flow, flowback, mask_fw, mask_bw, self.train_params["num_predicted_frames"])
