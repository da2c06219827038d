"""Does the forward of an NCHW x get cheaper when the relayout and the channels-last forward run chunk by chunk, so
that the channels-last copy of a chunk is still in L2 when the forward reads it?  (40 x 64 x 256 x 512; CUDA events.)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from c2m_b200 import _lib

dev = torch.device("cuda", 0)
N, C, H, W = 40, 64, 256, 512
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(N, C, H, W, device=dev, generator=g)
flow = torch.randn(N, 2, H, W, device=dev, generator=g) * 3
mask = torch.rand(N, 1, H, W, device=dev, generator=g)
xcl = torch.empty_like(x, memory_format=torch.channels_last)
out = torch.empty_like(xcl)
st = torch.cuda.current_stream().cuda_stream


def run(chunk):
    for n0 in range(0, N, chunk):
        n = min(chunk, N - n0)
        _lib.relayout(x[n0:].data_ptr(), xcl[n0:].data_ptr(), n, C, H, W, True, st)
        _lib.warp_blend_fwd(xcl[n0:].data_ptr(), flow[n0:].data_ptr(), mask[n0:].data_ptr(), None, out[n0:].data_ptr(),
                            n, C, H, W, n, xcl.stride(), out.stride(), 0, 0, st, None)


for chunk in (40, 20, 8, 4, 2, 1):
    for _ in range(3):
        run(chunk)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run(chunk)
    e1.record()
    torch.cuda.synchronize()
    print(f"chunk {chunk:3d} frames: relayout + forward {e0.elapsed_time(e1) / 5:.3f} ms")
