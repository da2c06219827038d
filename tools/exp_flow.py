"""Experiment (development tool): forward time vs flow pattern, to separate access-pattern effects from kernel structure."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import c2m_b200
from bench import synth, fwd_bytes

dev = torch.device("cuda", 0)
N, C, H, W = 40, 64, 256, 512
x, flow, mask, gout = synth(N, C, H, W, False, 1234, dev)
jj = torch.arange(W, device=dev, dtype=torch.float32).view(1, 1, W)
ii = torch.arange(H, device=dev, dtype=torch.float32).view(1, H, 1)
ident = torch.stack([((jj + 0.5) * (W - 1) / W - jj).expand(N, H, W), ((ii + 0.5) * (H - 1) / H - ii).expand(N, H, W)], 1).contiguous()
flows = {"synthetic": flow, "identity": ident, "zero": torch.zeros_like(flow), "shift8": ident + 8.0,
         "noise_only": ident + torch.randn_like(ident), "smooth_only": flow - (flow - synth(N, C, H, W, False, 1234, dev)[1]) }

def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

xl = x.contiguous(memory_format=torch.channels_last)
fb = fwd_bytes(N, C, H, W)
for name, f in flows.items():
    a = timeit(lambda: c2m_b200.warp_blend(x, f, mask))
    b = timeit(lambda: c2m_b200.warp_blend(xl, f, mask))
    print(f"{name:12s} nchw {a:7.3f} ms {fb/a/1e6:7.0f} GB/s | nhwc {b:7.3f} ms {fb/b/1e6:7.0f} GB/s", flush=True)
