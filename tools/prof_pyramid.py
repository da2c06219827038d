"""Per-level timing of the multi-scale sites (development tool): eager (host-bound for the small levels),
CUDA-graph replay (GPU time only) and the torch composition, N frames per level.
    python tools/prof_pyramid.py [--frames 40] [--iters 20] [--once]     (--once: 2 eager passes only, for ncu)"""
import argparse
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import c2m_b200  # noqa: E402
from bench import synth, fwd_bytes, bwd_bytes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=40)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--once", action="store_true")
ap.add_argument("--layout", default="nhwc")
ap.add_argument("--flags", default="0")
ap.add_argument("--graph-only", action="store_true")
ap.add_argument("--only-c", type=int, default=0, help="run only the levels with this channel count")
a = ap.parse_args()
dev = torch.device("cuda", 0)
LEVELS = [(32, 256, 512), (64, 128, 256), (128, 64, 128), (256, 32, 64),
          (64, 64, 128), (128, 32, 64), (256, 16, 32), (512, 8, 16), (3, 256, 512)]
N = a.frames


def torch_step(x, flow, mask, gout):
    n, _, h, w = flow.shape
    g0 = torch.zeros([n, 2, h, w])
    g0[:, 0] = torch.linspace(-1, 1, w).view(1, 1, w).expand(n, h, w)
    g0[:, 1] = torch.linspace(-1, 1, h).view(1, h, 1).expand(n, h, w)
    g0 = g0.to(x.device)
    nf = torch.cat([flow[:, 0:1] / ((w - 1.0) / 2.0), flow[:, 1:2] / ((h - 1.0) / 2.0)], dim=1)
    o = F.grid_sample(x, (g0 + nf).permute(0, 2, 3, 1), mode="bilinear", padding_mode="border",
                      align_corners=False) * mask
    torch.autograd.grad(o, [x, flow, mask], gout)


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3, (time.perf_counter() - t0) / iters * 1e6


for (c, h, w) in LEVELS:
    if a.only_c and c != a.only_c:
        continue
    x, flow, mask, gout = synth(N, c, h, w, False, 77, dev)
    if a.layout == "nhwc" and c % 4 == 0:
        x, gout = x.contiguous(memory_format=torch.channels_last), gout.contiguous(memory_format=torch.channels_last)
    x.requires_grad_(True), flow.requires_grad_(True), mask.requires_grad_(True)

    def step():
        o = c2m_b200.warp_blend(x, flow, mask, flags=int(a.flags, 0))
        return torch.autograd.grad(o, [x, flow, mask], gout)

    if a.once:
        step(), step()
        torch.cuda.synchronize()
        continue
    ev_us, wall_us = (0.0, 0.0) if a.graph_only else min(timed(step, a.iters) for _ in range(3))  # best of three
    # graph replay: the same launches without the host
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    gr_us, _ = timed(g.replay, a.iters)
    th_us = 0.0 if a.graph_only else timed(lambda: torch_step(x, flow, mask, gout), max(3, a.iters // 4))[0]
    by = fwd_bytes(N, c, h, w) + bwd_bytes(N, c, h, w)
    print(f"C={c:4d} {h:4d}x{w:<4d} N={N}: eager {ev_us:8.1f} us (host wall {wall_us:7.1f}), graph {gr_us:8.1f} us "
          f"= {by / gr_us / 1e3:7.1f} GB/s, torch composition {th_us:9.1f} us, bytes {by / 1e6:8.1f} MB", flush=True)
print("done")
