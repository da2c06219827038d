"""TEST INFRASTRUCTURE -- torch restatement of the reference warp path (not shipped, not measured
as product).  Device-agnostic: on CPU it is the "reference CPU path" (BASELINE.json configs[0]),
on CUDA it is the reference's own GPU arithmetic (ATen grid_sampler_2d) and therefore the primary
parity target for the sm_100a kernels.

Each function names the reference lines it restates.  The single intentional deviation is the
device fix: the reference moves its CPU-built base grid with ``.cuda(gpu_id)``
(/root/reference/src/utils/ops.py:202), which raises on CPU tensors (gpu_id == -1); here the grid
follows ``flow.device``.  The grid is still built on the CPU in float32 first, as the reference
does, so its bits are identical.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def base_grid(batch: int, rows: int, cols: int, device) -> torch.Tensor:
    """ops.py:196-202 ``get_grid``: channel 0 = linspace(-1,1,cols) along x, channel 1 =
    linspace(-1,1,rows) along y, float32, built on the CPU, then moved."""
    g = torch.zeros([batch, 2, rows, cols])
    xs = torch.linspace(-1, 1, cols) if cols > 1 else torch.Tensor([-1])
    ys = torch.linspace(-1, 1, rows) if rows > 1 else torch.Tensor([-1])
    g[:, 0] = xs.view(1, 1, cols).expand(batch, rows, cols)
    g[:, 1] = ys.view(1, rows, 1).expand(batch, rows, cols)
    return g.to(device)


def grid_sample_border(image: torch.Tensor, grid: torch.Tensor, mode: str = "bilinear") -> torch.Tensor:
    """ops.py:183-184: F.grid_sample, padding_mode='border', align_corners left at its default
    (False)."""
    return F.grid_sample(image, grid, mode=mode, padding_mode="border", align_corners=False)


def resample(image: torch.Tensor, flow: torch.Tensor, mode: str = "bilinear") -> torch.Tensor:
    """ops.py:187-193: pixel flow -> normalised offsets with the (size-1)/2 convention, added to
    the base grid, sampled with the align_corners=False convention (SURVEY.md section 0,
    quirk 1)."""
    b, c, h, w = image.size()
    grid = base_grid(b, h, w, flow.device)
    # tensor / python-float, exactly as the reference writes it (CUDA: reciprocal multiply,
    # CPU: true division -- SURVEY.md appendix A.3 step 2)
    nflow = torch.cat([flow[:, 0:1] / ((w - 1.0) / 2.0), flow[:, 1:2] / ((h - 1.0) / 2.0)], dim=1)
    final_grid = (grid + nflow).permute(0, 2, 3, 1)
    return grid_sample_border(image, final_grid, mode)


def deform_input(inp: torch.Tensor, optical_flow: torch.Tensor) -> torch.Tensor:
    """generator.py:80-86.  The reference unpacks the NCHW flow as if it were NHWC, so the size
    test compares (C, H) of the flow with (H, W) of the input and the flow goes through
    F.interpolate whenever that (mis-)comparison differs; values are not rescaled."""
    _, h_old, w_old, _ = optical_flow.shape
    _, _, h, w = inp.shape
    if h_old != h or w_old != w:
        optical_flow = F.interpolate(optical_flow, size=(h, w), mode="bilinear")
    return resample(inp, optical_flow)


def apply_optical(input_ref, optical_flow, occlusion_map=None):
    """generator.py:88-96: warp, then multiply by the (resized if needed) occlusion map."""
    warped = deform_input(input_ref, optical_flow)
    if occlusion_map is None:
        return warped
    if warped.shape[2] != occlusion_map.shape[2] or warped.shape[3] != occlusion_map.shape[3]:
        occlusion_map = F.interpolate(occlusion_map, size=warped.shape[2:], mode="bilinear")
    return warped * occlusion_map


def resize_flow(flow: torch.Tensor, new_shape) -> torch.Tensor:
    """utils.py:346-354: bilinear align_corners=True resize with the values rescaled by
    new/old."""
    _, _, h, w = flow.shape
    new_h, new_w = new_shape
    out = F.interpolate(flow, (new_h, new_w), mode="bilinear", align_corners=True)
    scale_h, scale_w = h / float(new_h), w / float(new_w)
    out[:, 0] /= scale_w
    out[:, 1] /= scale_h
    return out


def decoder_warp(app_features, sparse_motion, sparse_occlusion, num_frames: int):
    """motion_autoencoder.py:117-125 (one scale): repeat the appearance map T times folded into
    the batch, resize motion (resize_flow) and occlusion (bilinear), warp and multiply.
    ``sparse_motion`` [B,2,T,H,W], ``sparse_occlusion`` [B,1,T,H,W]."""
    rep = torch.cat(torch.unbind(app_features.unsqueeze(2).repeat(1, 1, num_frames, 1, 1), dim=2), dim=0)
    nh, nw = rep.shape[-2:]
    motion = resize_flow(torch.cat(torch.unbind(sparse_motion, 2), 0), [nh, nw])
    occ = F.interpolate(torch.cat(torch.unbind(sparse_occlusion, 2), 0), size=[nh, nw], mode="bilinear")
    return resample(rep, motion) * occ


def warp_blend(x, flow, mask=None, other=None):
    """The fused op's contract in reference terms: ``resample(x, flow) * mask`` (reference form,
    other=None), or the north-star blend ``m*warp + (1-m)*other`` (extension)."""
    w = resample(x, flow)
    if mask is None:
        return w
    if other is None:
        return w * mask
    return w * mask + (1.0 - mask) * other


def warped_l1(source: torch.Tensor, flows: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
    """The `warped` term of the training losses (src/losses/losses.py:219-222): the source frame warped by each of
    the T backward flows, stacked on a frame axis, L1 against the target frames (L1MaskedLoss without a mask,
    losses.py:184-189, i.e. F.l1_loss with mean reduction)."""
    frames = [resample(source, flows[:, :, t]).unsqueeze(2) for t in range(flows.shape[2])]
    return F.l1_loss(torch.cat(frames, dim=2), targets)


def affine_warp(affine_matrix: torch.Tensor, x: torch.Tensor, base_grid_nhw2: torch.Tensor):
    """src/modules/motion_estimator/dense_motion.py:161-168 ``DenseMotionNetwork.warp``: affine_grid (align_corners
    left at its default False), flow in pixels relative to the linspace base grid, grid_sample with its defaults
    (bilinear, zeros padding, align_corners False).  affine_matrix [2,3], x [1,C,h,w], base_grid [1,h,w,2]."""
    grid = F.affine_grid(affine_matrix.unsqueeze(0), x.size(), align_corners=False)
    b, _, h, w = x.size()
    flow = grid - base_grid_nhw2
    flow = torch.cat([flow[:, :, :, 0:1] * ((w - 1.0) / 2.0), flow[:, :, :, 1:2] * ((h - 1.0) / 2.0)], dim=-1)
    t_x = F.grid_sample(x, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    return t_x, flow.permute(0, 3, 1, 2)


def object_base_grid(h: int, w: int, device) -> torch.Tensor:
    """dense_motion.py:118-123: the [1,h,w,2] linspace grid of generate_sparse_motion (built on the CPU, moved)."""
    base = torch.zeros([1, h, w, 2])
    lx = torch.linspace(-1, 1, w) if w > 1 else torch.Tensor([-1])
    base[:, :, :, 0] = torch.ger(torch.ones(h), lx).expand_as(base[:, :, :, 0])
    ly = torch.linspace(-1, 1, h) if h > 1 else torch.Tensor([-1])
    base[:, :, :, 1] = torch.ger(ly, torch.ones(w)).expand_as(base[:, :, :, 1])
    return base.to(device)


def generate_sparse_motion(source_instance, inst_ids, batch_ids, thetas, num_frames: int):
    """dense_motion.py:94-152 with the tracking_gnn / dict arguments unpacked into plain tensors: source_instance
    [B,1,H,W]; inst_ids [n_obj] (long), batch_ids [n_obj] (long), thetas [n_obj,T,6].  Returns (sparse_motion_bw,
    sparse_motion_fw [B,2,T,H,W], sparse_motion_bin [B,1,T,H,W]) before the detach / occlusion-map lines (:149-158)."""
    B, _, h, w = source_instance.shape
    dev = source_instance.device
    bw = torch.zeros(B, 2, num_frames, h, w, device=dev)
    fw = torch.zeros(B, 2, num_frames, h, w, device=dev)
    bn = torch.zeros(B, 1, num_frames, h, w, device=dev)
    base = object_base_grid(h, w, dev)
    for idx, (inst_id, batch_id) in enumerate(zip(inst_ids.long(), batch_ids.long())):
        if inst_id == 0:
            continue
        obj_mask = (source_instance[batch_id] == torch.squeeze(inst_id)).float()
        for t in range(num_frames):
            warped_obj, obj_flow = affine_warp(thetas[idx, t].view(2, 3), obj_mask.unsqueeze(0), base)
            bw[batch_id, :, t, ...] = torch.where(warped_obj == 1, obj_flow, bw[batch_id, :, t, ...])
            fw[batch_id, :, t, ...] = torch.where(obj_mask == 1, obj_flow * -1, fw[batch_id, :, t, ...])
            bn[batch_id, :, t, ...] = torch.where(warped_obj == 1, warped_obj, bn[batch_id, :, t, ...])
    return bw, fw, bn


def flow_consistency(flow, flowback, mask_fw=None, mask_bw=None):
    """src/losses/losses.py:122-129 ``FlowConsistLoss._flowconsist`` on the folded [N,2,H,W] tensors."""
    if mask_fw is not None:
        nextloss = (mask_fw * torch.abs(resample(flowback, flow) + flow)).mean()
        prevloss = (mask_bw * torch.abs(resample(flow, flowback) + flowback)).mean()
    else:
        nextloss = torch.abs(resample(flowback, flow) + flow).mean()
        prevloss = torch.abs(resample(flow, flowback) + flowback).mean()
    return prevloss + nextloss


def flow_consistency_loss(flow, flowback, mask_fw=None, mask_bw=None, num_predicted_frames: int = 5):
    """losses.py:131-141 ``FlowConsistLoss.forward``: fold the frame axis (t-major) and scale by the frame count.
    flow / flowback [B,2,T,H,W], masks [B,1,T,H,W] or None."""
    fold = lambda t: torch.cat(torch.unbind(t, dim=2), dim=0)  # noqa: E731
    if mask_bw is not None:
        v = flow_consistency(fold(flow), fold(flowback), fold(mask_fw), fold(mask_bw))
    else:
        v = flow_consistency(fold(flow), fold(flowback))
    return v * num_predicted_frames
