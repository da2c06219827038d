"""Host-side mirror of the affine-grid object warp of the reference's dense-motion network
(/root/reference/src/modules/motion_estimator/dense_motion.py):

    affine_warp(affine_matrix, x, base_grid=None)               DenseMotionNetwork.warp, :161-168 (batched)
    sparse_motion(source_instance, inst_ids, batch_ids, thetas)  the objects x T loop of :94-152 in ONE launch
    generate_sparse_motion(self, tracking_gnn, sparse_motion_dict, source_instance, use_gt=False)
                                                                drop-in for the reference method, :94-159

The reference calls `warp` objects x T times from Python (each call: affine_grid, sub, two muls, cat, grid_sample, then
three torch.where merges): ~15 launches per (object, frame).  Here the whole loop is one kernel; the object masks
`(instance == id).float()` are formed on the fly and never materialised.
"""
from __future__ import annotations

import torch

from . import _lib
from .ops import get_occlusion_map


def _f32_cuda(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"c2m_b200.motion: `{what}` must be a CUDA tensor (no CPU fallback)")
    return t.detach().to(torch.float32).contiguous()


def affine_warp(affine_matrix: torch.Tensor, x: torch.Tensor, base_grid=None):
    """`DenseMotionNetwork.warp(affine_matrix, x, base_grid)` (a staticmethod in the reference), batched.

    affine_matrix [2,3] (the reference's call) or [K,2,3]; x [1,C,h,w], [K,C,h,w], or [Kx,C,h,w] with Kx dividing K
    (theta k samples image k % Kx).  Returns (t_x [K,C,h,w], flow [K,2,h,w]) -- for K == 1 exactly the reference's
    shapes.  `base_grid` is accepted for signature compatibility only: the kernel forms the reference's linspace
    grid (dense_motion.py:118-123) in registers, bit for bit.  Forward only (the reference detaches the flows and the
    warped masks are piecewise constant in theta)."""
    theta = _f32_cuda(affine_matrix, "affine_matrix").reshape(-1, 2, 3)
    xx = _f32_cuda(x, "x")
    if xx.dim() != 4:
        raise ValueError(f"x must be [K,C,h,w], got {tuple(x.shape)}")
    K, Kx = theta.shape[0], xx.shape[0]
    if Kx == 0 or K % Kx != 0:
        raise ValueError(f"x batch {Kx} must divide the number of affine matrices {K}")
    _, C, H, W = xx.shape
    t_x = torch.empty((K, C, H, W), dtype=torch.float32, device=xx.device)
    flow = torch.empty((K, 2, H, W), dtype=torch.float32, device=xx.device)
    with torch.cuda.device(xx.device):
        _lib.affine_warp(theta.data_ptr(), xx.data_ptr(), None, t_x.data_ptr(), flow.data_ptr(), K, Kx, C, H, W,
                         torch.cuda.current_stream().cuda_stream)
    return t_x, flow


def sparse_motion(source_instance: torch.Tensor, inst_ids: torch.Tensor, batch_ids: torch.Tensor, thetas: torch.Tensor,
                  want_fw: bool = True):
    """The object loop of generate_sparse_motion (dense_motion.py:124-148) as one launch.

    source_instance [B,1,H,W]; inst_ids [n_obj] (0 = skipped); batch_ids [n_obj]; thetas [n_obj,T,6] (or [n_obj,T,2,3]).
    Returns (sparse_motion_bw [B,2,T,H,W], sparse_motion_fw or None, sparse_motion_bin [B,1,T,H,W])."""
    inst = _f32_cuda(source_instance, "source_instance")
    if inst.dim() != 4 or inst.shape[1] != 1:
        raise ValueError(f"source_instance must be [B,1,H,W], got {tuple(source_instance.shape)}")
    B, _, H, W = inst.shape
    n_obj = int(inst_ids.shape[0])
    th = _f32_cuda(thetas, "thetas")
    th = th.reshape(n_obj, -1, 6) if n_obj > 0 else th.reshape(0, th.shape[1] if th.dim() > 1 else 0, 6)
    T = th.shape[1]
    ids = inst_ids.detach().to(device=inst.device, dtype=torch.float32).contiguous()
    bat = batch_ids.detach().to(device=inst.device, dtype=torch.int32).contiguous()
    dev = inst.device
    bw = torch.empty((B, 2, T, H, W), dtype=torch.float32, device=dev)
    fw = torch.empty((B, 2, T, H, W), dtype=torch.float32, device=dev) if want_fw else None
    bn = torch.empty((B, 1, T, H, W), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.sparse_motion(inst.data_ptr(), ids.data_ptr(), bat.data_ptr(), th.data_ptr(), bw.data_ptr(),
                           None if fw is None else fw.data_ptr(), bn.data_ptr(), B, T, H, W, n_obj,
                           torch.cuda.current_stream().cuda_stream)
    return bw, fw, bn


def clip_mask(mask: torch.Tensor) -> torch.Tensor:
    """dense_motion.py:154-158."""
    return (mask > 0.5).to(mask.dtype)


def generate_sparse_motion(self, tracking_gnn, sparse_motion_dict, source_instance, use_gt=False):
    """Drop-in for `DenseMotionNetwork.generate_sparse_motion` (dense_motion.py:94-159): same arguments, same
    dictionary.  One launch for the object loop, one for each direction's occlusion maps (the T per-frame calls of
    utils.get_occlusion_map folded into the batch axis)."""
    T = self.train_params["num_predicted_frames"]
    ids = tracking_gnn.source_frames_nodes_instance_ids[:, -1]
    batch = tracking_gnn.batch
    if use_gt:
        thetas = tracking_gnn.targets_theta.reshape(ids.shape[0], -1, 6)[:, :T]
    else:
        thetas = torch.stack([sparse_motion_dict[f"theta_{t}"].reshape(ids.shape[0], 6) for t in range(T)], dim=1)
    bw, fw, bn = sparse_motion(source_instance, ids, batch, thetas, want_fw=True)
    out = {"sparse_motion_bw": bw}
    if self.train_params["use_fw_of"]:
        out["sparse_motion_fw"] = fw
    out["sparse_motion_bin"] = bn
    B, _, _, H, W = bw.shape

    def occ(flows):  # [B,2,T,H,W] -> per-frame occlusion maps, frames folded into the batch (t-major), then unfolded
        folded = flows.permute(2, 0, 1, 3, 4).reshape(T * B, 2, H, W)
        return clip_mask(get_occlusion_map(folded)).view(T, B, 1, H, W).permute(1, 2, 0, 3, 4).contiguous()

    out["sparse_occ_bw"] = occ(fw)
    out["sparse_occ_fw"] = occ(bw)
    return out
