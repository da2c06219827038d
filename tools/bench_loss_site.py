"""The `warped` loss term (losses.py:219-222) at the configs[1] shape, B=8 clips of T=5 frames, C=3, 256x512
(development tool): fused kernel vs T calls of the library's resample + l1_loss vs the reference's torch composition
(CPU-built grid copied to the device every call, as ops.py:187-202 does)."""
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import c2m_b200  # noqa: E402

dev = torch.device("cuda", 0)
B, C, T, H, W = 8, 3, 5, 256, 512
torch.manual_seed(0)
source = torch.randn(B, C, H, W, device=dev)
targets = torch.randn(B, C, T, H, W, device=dev)
# flows as in the headline benchmark: 8 px low-frequency field + 1 px noise
ii = torch.arange(H, device=dev, dtype=torch.float32).view(1, 1, H, 1)
jj = torch.arange(W, device=dev, dtype=torch.float32).view(1, 1, 1, W)
fx = 8.0 * torch.sin(6.2831853 * ii / (H / 2.0)) * torch.cos(6.2831853 * jj / (W / 2.0))
fy = 8.0 * torch.cos(6.2831853 * ii / (H / 2.0)) * torch.sin(6.2831853 * jj / (W / 2.0))
flows = (torch.stack([fx.expand(B, T, H, W), fy.expand(B, T, H, W)], 1) + torch.randn(B, 2, T, H, W, device=dev))
flows = flows.requires_grad_(True)


def ref_resample(image, flow):
    b, _, h, w = image.shape
    g0 = torch.zeros([b, 2, h, w])
    g0[:, 0] = torch.linspace(-1, 1, w).view(1, 1, w).expand(b, h, w)
    g0[:, 1] = torch.linspace(-1, 1, h).view(1, h, 1).expand(b, h, w)
    g0 = g0.to(image.device)
    nf = torch.cat([flow[:, 0:1] / ((w - 1.0) / 2.0), flow[:, 1:2] / ((h - 1.0) / 2.0)], dim=1)
    return F.grid_sample(image, (g0 + nf).permute(0, 2, 3, 1), mode="bilinear", padding_mode="border", align_corners=False)


def fused():
    torch.autograd.grad(c2m_b200.warped_l1_loss(source, flows, targets), [flows])


def ours_loop():
    w = torch.cat([c2m_b200.resample(source, flows[:, :, t]).unsqueeze(2) for t in range(T)], 2)
    torch.autograd.grad(F.l1_loss(w, targets), [flows])


def torch_loop():
    w = torch.cat([ref_resample(source, flows[:, :, t]).unsqueeze(2) for t in range(T)], 2)
    torch.autograd.grad(F.l1_loss(w, targets), [flows])


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    return best


def graphed(fn):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return timed(g.replay)


by = 4 * B * T * H * W * (2 * C + 2 + 2) + 4 * B * C * H * W  # fwd+bwd: targets, flows twice; gflows; source
print(f"loss site B={B} T={T} C={C} {H}x{W} forward+backward")
a = timed(fused)
print(f"  fused warped_l1_loss              {a:8.3f} ms   ({by / a / 1e6:7.1f} GB/s of {by / 1e6:.0f} MB algorithmic)")
a = graphed(fused)
print(f"  fused, CUDA-graph replay          {a:8.3f} ms   ({by / a / 1e6:7.1f} GB/s)")
print(f"  T x c2m_b200.resample + l1_loss   {timed(ours_loop):8.3f} ms")
print(f"  reference torch composition       {timed(torch_loop, 5):8.3f} ms")
# (a process that exits right after a backward can meet PyTorch's autograd worker thread still releasing the last
# graph's tensors while the interpreter finalises -- PyEval_AcquireThread -> pthread_exit -> std::terminate, seen on
# the GPU boxes with a native backtrace, tools/diag/term_trace.cpp; give that thread a moment)
torch.cuda.synchronize()
time.sleep(0.2)
