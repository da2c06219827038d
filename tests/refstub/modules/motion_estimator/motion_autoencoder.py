"""Stand-in for src/modules/motion_estimator/motion_autoencoder.py: binds `resample` at import time (:8); `warp_block`
is one scale of the decoder's warp (:115-125) in this repo's own wording."""
import torch
import torch.nn.functional as F

from utils import resample


def _fold_frames(t):  # [B, C, T, H, W] -> [T*B, C, H, W], frame-major
    return t.transpose(1, 2).transpose(0, 1).reshape(-1, t.shape[1], *t.shape[3:])


def warp_block(app_features, sparse_motion, sparse_occlusion, num_frames):
    size = tuple(app_features.shape[-2:])
    motion = F.interpolate(_fold_frames(sparse_motion), size, mode="bilinear", align_corners=True)
    sx, sy = size[1] / float(sparse_motion.shape[-1]), size[0] / float(sparse_motion.shape[-2])
    motion = torch.stack([motion[:, 0] * sx, motion[:, 1] * sy], 1)
    occ = F.interpolate(_fold_frames(sparse_occlusion), size=size, mode="bilinear")
    return resample(app_features.repeat(num_frames, 1, 1, 1), motion) * occ
