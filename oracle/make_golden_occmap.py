"""TEST INFRASTRUCTURE -- generates tests/golden/occmap/*.npz by running the UNMODIFIED reference functions
utils.ops.get_occlusion_map / get_corresponding_map (/root/reference/src/utils/ops.py:205-275) on the CPU of the
build container (they are plain torch ops and run on CPU tensors as shipped; ``imageio`` is stubbed as in
oracle/make_golden.py).

Run:  python oracle/make_golden_occmap.py      (needs /root/reference; the GPU box never runs this)
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/src"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "occmap")


def main():
    sys.modules.setdefault("imageio", types.ModuleType("imageio"))
    sys.path.insert(0, REF)
    import utils.ops as ref_ops
    os.makedirs(OUT, exist_ok=True)
    g = torch.Generator().manual_seed(20261019)
    cases = {}
    cases["zero_flow_6x9"] = torch.zeros(2, 2, 6, 9)
    cases["noise_8x16"] = torch.randn(2, 2, 8, 16, generator=g) * 2
    f = torch.zeros(1, 2, 8, 12)
    f[:, 0], f[:, 1] = 3.0, -2.0
    cases["integer_shift_8x12"] = f
    cases["half_pixel_5x7"] = torch.full((1, 2, 5, 7), 0.5)
    cases["out_of_bounds_8x16"] = torch.randn(1, 2, 8, 16, generator=g) * 12
    ii = torch.arange(16, dtype=torch.float32).view(1, 16, 1)
    jj = torch.arange(32, dtype=torch.float32).view(1, 1, 32)
    cases["converging_16x32"] = torch.stack([((16 - jj) * 0.5).expand(2, 16, 32), ((8 - ii) * 0.5).expand(2, 16, 32)], 1) \
        + 0.3 * torch.randn(2, 2, 16, 32, generator=g)
    cases["smooth_64x128"] = torch.stack([
        (4 * torch.sin(6.2831853 * ii[:, :1].new_tensor(range(64)).view(1, 64, 1) / 32) * torch.ones(1, 1, 128)).expand(1, 64, 128),
        (4 * torch.cos(6.2831853 * torch.arange(128, dtype=torch.float32).view(1, 1, 128) / 64) * torch.ones(1, 64, 1)).expand(1, 64, 128)], 1) \
        + torch.randn(1, 2, 64, 128, generator=g)
    for name, flow in cases.items():
        flow = flow.contiguous()
        occ = ref_ops.get_occlusion_map(flow)
        base = ref_ops.mesh_grid(flow.shape[0], flow.shape[2], flow.shape[3]).type_as(flow)
        corr = ref_ops.get_corresponding_map(base + flow)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), flow=flow.numpy(), occ=occ.numpy(), corr=corr.numpy())
        print(name, tuple(flow.shape), float(occ.min()), float(occ.max()), float(corr.max()))


if __name__ == "__main__":
    main()
