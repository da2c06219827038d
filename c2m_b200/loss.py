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

__all__ = ["warped_l1_loss", "WarpedL1Function"]


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
