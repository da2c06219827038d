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


# Second experiment: the same chunks on TWO streams -- the relayout of chunk k+1 runs next to the forward of chunk k
# (events order them), so the launch tails of one stream are covered by the other while the channels-last copy of a
# chunk is read back out of L2.
s_a, s_b = torch.cuda.Stream(), torch.cuda.Stream()


def run2(chunk):
    cur = torch.cuda.current_stream()
    s_a.wait_stream(cur)
    s_b.wait_stream(cur)
    evs = []
    for n0 in range(0, N, chunk):
        n = min(chunk, N - n0)
        _lib.relayout(x[n0:].data_ptr(), xcl[n0:].data_ptr(), n, C, H, W, True, s_a.cuda_stream)
        ev = torch.cuda.Event()
        ev.record(s_a)
        evs.append(ev)
    for k, n0 in enumerate(range(0, N, chunk)):
        n = min(chunk, N - n0)
        s_b.wait_event(evs[k])
        _lib.warp_blend_fwd(xcl[n0:].data_ptr(), flow[n0:].data_ptr(), mask[n0:].data_ptr(), None, out[n0:].data_ptr(),
                            n, C, H, W, n, xcl.stride(), out.stride(), 0, 0, s_b.cuda_stream, None)
    cur.wait_stream(s_a)
    cur.wait_stream(s_b)


ref = out.clone()
for chunk in (40, 8, 4, 2, 1):
    for _ in range(3):
        run2(chunk)
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run2(chunk)
    e1.record()
    torch.cuda.synchronize()
    print(f"two streams, chunk {chunk:3d} frames: relayout + forward {e0.elapsed_time(e1) / 5:.3f} ms")


# Third experiment: the two-stream schedule captured in a CUDA graph (no host time between the launches: the small
# chunks above are host-bound) -- the cleanest answer to "does reading the channels-last copy out of L2 pay".
for chunk in (8, 4, 2, 1):
    g = torch.cuda.CUDAGraph()
    cap = torch.cuda.Stream()
    cap.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(cap):
        run2(chunk)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=cap):
            run2(chunk)
    torch.cuda.current_stream().wait_stream(cap)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"two streams in a CUDA graph, chunk {chunk:3d} frames: relayout + forward {e0.elapsed_time(e1) / 5:.3f} ms")
