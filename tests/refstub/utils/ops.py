"""Stand-in for src/utils/ops.py: the names the reference defines there, bodies = the oracle's restatements."""
import torch

from oracle import reference_torch as _rt
from oracle import occmap_numpy as _on  # noqa: F401  (kept importable: the occlusion-map oracle)

__all__ = ["resample", "grid_sample", "get_grid", "get_occlusion_map", "get_corresponding_map"]


def grid_sample(input1, input2, mode="bilinear"):
    return _rt.grid_sample_border(input1, input2, mode)


def resample(image, flow, mode="bilinear"):
    return _rt.resample(image, flow, mode)


def get_grid(batchsize, rows, cols, gpu_id=0):
    return _rt.base_grid(batchsize, rows, cols, torch.device("cuda", gpu_id) if gpu_id >= 0 else "cpu")


def get_corresponding_map(data):
    return torch.from_numpy(_on.corresponding_map(data.detach().cpu().numpy())).to(data.device)


def get_occlusion_map(flow):
    return torch.from_numpy(_on.occlusion_map(flow.detach().cpu().numpy())).to(flow.device)
