"""Stand-in for src/modules/motion_estimator/motion_autoencoder.py (the warp block, :115-125)."""
import torch
import torch.nn.functional as F

from utils import resample


def resize_flow(flow, new_shape):
    _, _, h, w = flow.shape
    new_h, new_w = new_shape
    out = F.interpolate(flow, (new_h, new_w), mode="bilinear", align_corners=True)
    out[:, 0] /= w / float(new_w)
    out[:, 1] /= h / float(new_h)
    return out


def warp_block(app_features, sparse_motion, sparse_occlusion, num_frames):
    rep = torch.cat(torch.unbind(app_features.unsqueeze(2).repeat(1, 1, num_frames, 1, 1), dim=2), dim=0)
    nh, nw = rep.shape[-2:]
    motion = resize_flow(torch.cat(torch.unbind(sparse_motion, 2), 0), [nh, nw])
    occ = F.interpolate(torch.cat(torch.unbind(sparse_occlusion, 2), 0), size=[nh, nw], mode="bilinear")
    return resample(rep, motion) * occ
