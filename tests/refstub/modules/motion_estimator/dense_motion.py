"""Stand-in for src/modules/motion_estimator/dense_motion.py (`warp` :161-168, `generate_sparse_motion` :94-159)."""
import torch
from torch import nn

import utils
from oracle import reference_torch as _rt


class DenseMotionNetwork(nn.Module):
    def __init__(self, train_params):
        super().__init__()
        self.train_params = train_params

    @staticmethod
    def clip_mask(mask):
        return torch.where(mask > 0.5, torch.ones_like(mask), torch.zeros_like(mask))

    @staticmethod
    def warp(affine_matrix, x, base_grid):
        return _rt.affine_warp(affine_matrix, x, base_grid)

    def generate_sparse_motion(self, tracking_gnn, sparse_motion_dict, source_instance, use_gt=False):
        T = self.train_params["num_predicted_frames"]
        ids = tracking_gnn.source_frames_nodes_instance_ids[:, -1]
        thetas = tracking_gnn.targets_theta if use_gt else torch.stack(
            [sparse_motion_dict[f"theta_{t}"] for t in range(T)], 1)
        bw, fw, bn = _rt.generate_sparse_motion(source_instance, ids, tracking_gnn.batch, thetas, T)
        out = {"sparse_motion_bw": bw.detach()}
        if self.train_params["use_fw_of"]:
            out["sparse_motion_fw"] = fw.detach()
        out["sparse_motion_bin"] = bn
        out["sparse_occ_bw"] = torch.cat([self.clip_mask(utils.get_occlusion_map(fw[:, :, i])).unsqueeze(2)
                                          for i in range(T)], 2)
        out["sparse_occ_fw"] = torch.cat([self.clip_mask(utils.get_occlusion_map(bw[:, :, i])).unsqueeze(2)
                                          for i in range(T)], 2)
        return out
