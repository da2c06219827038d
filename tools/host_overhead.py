"""Where the host time of one small fwd+bwd call goes (development tool): raw C-ABI calls, the autograd
Function, torch.autograd.grad.  Smallest decoder level (N=40, C=512, 8x16), channels-last."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import c2m_b200  # noqa: E402
from c2m_b200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)
N, C, H, W = 40, 512, 8, 16
x = torch.randn(N, C, H, W, device=dev).contiguous(memory_format=torch.channels_last)
flow = torch.randn(N, 2, H, W, device=dev)
mask = torch.rand(N, 1, H, W, device=dev)
gout = torch.randn(N, C, H, W, device=dev).contiguous(memory_format=torch.channels_last)
out = torch.empty_like(x)
gx, gf, gm = torch.empty_like(x), torch.empty_like(flow), torch.empty_like(mask)
wsb = _lib.bwd_workspace_bytes(N, C, H, W, N, True, 0)
ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream


def wall(fn, n=300):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t0) / n * 1e6)
    return best


def raw_fwd():
    _lib.warp_blend_fwd(x.data_ptr(), flow.data_ptr(), mask.data_ptr(), None, out.data_ptr(), N, C, H, W, N,
                        x.stride(), out.stride(), 0, 0, st)


def raw_bwd():
    _lib.warp_blend_bwd(x.data_ptr(), flow.data_ptr(), mask.data_ptr(), None, gout.data_ptr(), gx.data_ptr(),
                        gf.data_ptr(), gm.data_ptr(), None, N, C, H, W, N, x.stride(), gout.stride(), 0, 0,
                        ws.data_ptr(), wsb, st)


xr, fr, mr = x.clone().requires_grad_(True), flow.clone().requires_grad_(True), mask.clone().requires_grad_(True)


def fn_nograd():
    with torch.no_grad():
        c2m_b200.warp_blend(x, flow, mask)


def fn_fwd():
    c2m_b200.warp_blend(xr, fr, mr)


def fn_both():
    o = c2m_b200.warp_blend(xr, fr, mr)
    torch.autograd.grad(o, [xr, fr, mr], gout)


def torch_allocs():
    torch.empty_like(x), torch.empty_like(flow), torch.empty_like(mask), torch.empty(wsb, dtype=torch.uint8, device=dev)


print(f"raw c2m_warp_blend_fwd (ctypes)          {wall(raw_fwd):7.1f} us   (GPU-bound if the kernels take longer)")
print(f"raw c2m_warp_blend_bwd (ctypes)          {wall(raw_bwd):7.1f} us")
print(f"4 x torch.empty (outputs + workspace)    {wall(torch_allocs):7.1f} us")
print(f"warp_blend under no_grad                 {wall(fn_nograd):7.1f} us")
print(f"warp_blend recording the graph           {wall(fn_fwd):7.1f} us")
print(f"warp_blend + torch.autograd.grad         {wall(fn_both):7.1f} us")
