"""GPU probe (development tool, not product): settles the coordinate-arithmetic variant against the
reference's CUDA path and times kernel variants + the reference composition on the same device."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import c2m_b200  # noqa: E402
from c2m_b200 import _lib  # noqa: E402
from oracle import reference_torch as rt  # noqa: E402
from bench import synth, fwd_bytes, bwd_bytes  # noqa: E402

dev = torch.device("cuda", 0)
res = {}


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


# 1. coordinate variants
coords = {}
for (H, W) in [(256, 512), (256, 832), (32, 104), (16, 52), (8, 26), (128, 256), (64, 208), (33, 77), (1024, 2048)]:
    x, flow, mask, gout = synth(2, 4, H, W, False, 1, dev)
    ref = rt.warp_blend(x, flow, mask)
    row = {}
    for name, fl in [("recip+fma", 0), ("truediv+fma", _lib.FLAG_TRUE_DIV), ("recip+nofma", _lib.FLAG_NO_FMA),
                     ("truediv+nofma", _lib.FLAG_TRUE_DIV | _lib.FLAG_NO_FMA)]:
        out = c2m_b200.warp_blend(x, flow, mask, flags=fl | _lib.FLAG_FORCE_GENERIC)
        row[name] = rel(out, ref)
    coords[f"{H}x{W}"] = row
    print("coords", H, W, {k: f"{v:.2e}" for k, v in row.items()}, flush=True)
res["coords"] = coords

# 2. timings at the headline shape
N, C, H, W = 40, 64, 256, 512
x, flow, mask, gout = synth(N, C, H, W, False, 1234, dev)
fb, bb = fwd_bytes(N, C, H, W), bwd_bytes(N, C, H, W)
tim = {}
for v in range(6):
    ms = timeit(lambda: c2m_b200.warp_blend(x, flow, mask, flags=v << 24))
    tim[f"fwd_nchw_v{v}"] = (ms, fb / ms / 1e6)
ms = timeit(lambda: c2m_b200.warp_blend(x, flow, mask, flags=_lib.FLAG_NO_TMA))
tim["fwd_nchw_no_tma"] = (ms, fb / ms / 1e6)
ms = timeit(lambda: c2m_b200.warp_blend(x, flow, mask, flags=_lib.FLAG_FORCE_GENERIC))
tim["fwd_generic"] = (ms, fb / ms / 1e6)
xl = x.contiguous(memory_format=torch.channels_last)
gl = gout.contiguous(memory_format=torch.channels_last)
ms = timeit(lambda: c2m_b200.warp_blend(xl, flow, mask))
tim["fwd_nhwc"] = (ms, fb / ms / 1e6)
ms = timeit(lambda: c2m_b200.warp_blend(xl, flow, mask, flags=_lib.FLAG_NO_TMA))
tim["fwd_nhwc_no_tma"] = (ms, fb / ms / 1e6)


def fb_run(xx, gg, need=(True, True, True), **kw):
    xr = xx.detach().requires_grad_(need[0])
    fr = flow.detach().requires_grad_(need[1])
    mr = mask.detach().requires_grad_(need[2])
    out = c2m_b200.warp_blend(xr, fr, mr, **kw)
    torch.autograd.grad(out, [t for t in (xr, fr, mr) if t.requires_grad], gg)


ms = timeit(lambda: fb_run(x, gout), iters=5)
tim["fwdbwd_nchw"] = (ms, (fb + bb) / ms / 1e6)
ms = timeit(lambda: fb_run(xl, gl), iters=5)
tim["fwdbwd_nhwc"] = (ms, (fb + bb) / ms / 1e6)
ms = timeit(lambda: fb_run(x, gout, need=(True, False, False)), iters=5)
tim["fwdbwd_nchw_gx_only"] = (ms, 0)
ms = timeit(lambda: fb_run(x, gout, need=(False, True, True)), iters=5)
tim["fwdbwd_nchw_gflow_gmask_only"] = (ms, 0)
ms = timeit(lambda: fb_run(x, gout, deterministic=True), iters=3)
tim["fwdbwd_nchw_det"] = (ms, (fb + bb) / ms / 1e6)


# reference composition on the same GPU
def ref_fb():
    xr = x.detach().requires_grad_(True)
    fr = flow.detach().requires_grad_(True)
    mr = mask.detach().requires_grad_(True)
    out = rt.warp_blend(xr, fr, mr)
    torch.autograd.grad(out, [xr, fr, mr], gout)


ms = timeit(lambda: rt.warp_blend(x, flow, mask), iters=5)
tim["ref_torch_cuda_fwd"] = (ms, fb / ms / 1e6)
ms = timeit(ref_fb, iters=5)
tim["ref_torch_cuda_fwdbwd"] = (ms, (fb + bb) / ms / 1e6)
cp = torch.empty_like(x)
ms = timeit(lambda: cp.copy_(x))
tim["copy_1.34GB"] = (ms, 2 * x.numel() * 4 / ms / 1e6)
ms = timeit(lambda: cp.zero_())
tim["memset_1.34GB"] = (ms, x.numel() * 4 / ms / 1e6)
for k, (ms, gbs) in tim.items():
    print(f"{k:34s} {ms:9.3f} ms  {gbs:9.1f} GB/s", flush=True)
res["timings"] = {k: {"ms": v[0], "gbs": v[1]} for k, v in tim.items()}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)
