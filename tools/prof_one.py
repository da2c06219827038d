"""Profiling driver (development tool): a few launches of one configuration, for ncu."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import c2m_b200  # noqa: E402
from bench import WORKLOADS, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cityscapes_256x512_c64")
ap.add_argument("--frames", type=int, default=8)
ap.add_argument("--layout", default="nchw")
ap.add_argument("--flags", default="0")
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--fwd-only", action="store_true")
ap.add_argument("--deterministic", action="store_true")
a = ap.parse_args()
_, C, H, W, oob = WORKLOADS[a.workload]
dev = torch.device("cuda", 0)
x, flow, mask, gout = synth(a.frames, C, H, W, oob, 1234, dev)
if a.layout == "nhwc":
    x = x.contiguous(memory_format=torch.channels_last)
    gout = gout.contiguous(memory_format=torch.channels_last)
x.requires_grad_(True)
flow.requires_grad_(True)
mask.requires_grad_(True)
for _ in range(a.iters):
    out = c2m_b200.warp_blend(x, flow, mask, flags=int(a.flags, 0), deterministic=a.deterministic)
    if not a.fwd_only:
        torch.autograd.grad(out, [x, flow, mask], gout)
torch.cuda.synchronize()
print("done")
