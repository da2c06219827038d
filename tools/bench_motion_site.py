"""Timing of the round-2 call sites (development tool): the fused flow / mask resize at the generator and decoder
sites, the object-warp loop of generate_sparse_motion, and the flow-consistency loss -- this library against the
reference's torch composition (oracle.reference_torch) on the same GPU.  Cityscapes shapes: 8 clips x 5 frames."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import c2m_b200  # noqa: E402
from c2m_b200 import _lib  # noqa: E402
from oracle import reference_torch as rt  # noqa: E402
from oracle.make_golden_motion import scene  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def graphed(fn, iters=20):
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream(dev).wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return timed(g.replay, iters)


def launches(fn):
    n0 = _lib.launch_count()
    fn()
    torch.cuda.synchronize()
    return _lib.launch_count() - n0


B, T, H, W = 8, 5, 256, 512
N = B * T
print(f"== generator site (generator.py:88-96): apply_optical, {N} frames, feature map 256 x {H // 8} x {W // 8}, "
      f"flow / occlusion at {H} x {W}, forward + backward (grad input, flow, mask)")
feat = torch.randn(N, 256, H // 8, W // 8, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
flow = (torch.randn(N, 2, H, W, device=dev) * 4).requires_grad_(True)
occ = torch.rand(N, 1, H, W, device=dev).requires_grad_(True)
gout = torch.randn(N, 256, H // 8, W // 8, device=dev).contiguous(memory_format=torch.channels_last)


def site_ours():
    o = c2m_b200.apply_optical(None, feat, flow, occ)
    torch.autograd.grad(o, [feat, flow, occ], gout)


def site_unfused():  # round 1: torch's F.interpolate around this library's warp
    import torch.nn.functional as F
    f = F.interpolate(flow, size=feat.shape[2:], mode="bilinear")
    m = F.interpolate(occ, size=feat.shape[2:], mode="bilinear")
    o = c2m_b200.warp_blend(feat, f, m)
    torch.autograd.grad(o, [feat, flow, occ], gout)


def site_ref():
    o = rt.apply_optical(feat, flow, occ)
    torch.autograd.grad(o, [feat, flow, occ], gout)


print(f"  fused (resize inside the kernels)   {timed(site_ours):8.3f} ms eager   {graphed(site_ours):8.3f} ms graph   "
      f"{launches(site_ours)} library launches")
print(f"  F.interpolate + warp_blend          {timed(site_unfused):8.3f} ms eager   {graphed(site_unfused):8.3f} ms graph")
print(f"  reference torch composition         {timed(site_ref, 5):8.3f} ms eager")

print(f"== decoder sites (motion_autoencoder.py:115-133): 4 scales, B={B}, T={T}, sparse motion / occlusion at {H // 2} x {W // 2}, "
      "forward + backward (grad appearance features)")
motion = torch.randn(B, 2, T, H // 2, W // 2, device=dev) * 4
socc = torch.rand(B, 1, T, H // 2, W // 2, device=dev)
levels = [(512, 8, 16), (256, 16, 32), (128, 32, 64), (64, 64, 128)]
apps = [torch.randn(B, c, h, w, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True) for c, h, w in levels]
gouts = [torch.randn(B * T, c, h, w, device=dev).contiguous(memory_format=torch.channels_last) for c, h, w in levels]


def dec_ours():
    for a, g in zip(apps, gouts):
        torch.autograd.grad(c2m_b200.decoder_warp(a, motion, socc, T), [a], g)


def dec_unfused():
    import torch.nn.functional as F
    for a, g in zip(apps, gouts):
        nh, nw = a.shape[-2:]
        mo = c2m_b200.resize_flow(torch.cat(torch.unbind(motion, 2), 0), [nh, nw])
        oc = F.interpolate(torch.cat(torch.unbind(socc, 2), 0), size=[nh, nw], mode="bilinear")
        torch.autograd.grad(c2m_b200.warp_blend(a, mo, oc), [a], g)


def dec_ref():
    for a, g in zip(apps, gouts):
        torch.autograd.grad(rt.decoder_warp(a, motion, socc, T), [a], g)


print(f"  fused                               {timed(dec_ours):8.3f} ms eager   {graphed(dec_ours):8.3f} ms graph   "
      f"{launches(dec_ours)} library launches")
print(f"  resize_flow + interpolate + warp    {timed(dec_unfused):8.3f} ms eager   {graphed(dec_unfused):8.3f} ms graph")
print(f"  reference torch composition         {timed(dec_ref, 5):8.3f} ms eager")

print("== generate_sparse_motion (dense_motion.py:94-152): 3 images of 128 x 256, 12 objects each, 5 frames")
inst, ids, batch, thetas = (t.to(dev) for t in scene(torch.Generator().manual_seed(5), 3, 128, 256, 12, 5))
t_ours = timed(lambda: c2m_b200.sparse_motion(inst, ids, batch, thetas))
t_ref = timed(lambda: rt.generate_sparse_motion(inst, ids, batch, thetas, 5), 3)
print(f"  one kernel                          {t_ours:8.3f} ms   ({launches(lambda: c2m_b200.sparse_motion(inst, ids, batch, thetas))} launch)")
print(f"  reference loop (objects x T)        {t_ref:8.3f} ms")

print(f"== flow-consistency loss (losses.py:115-141): B={B}, T={T}, {H} x {W}, masked, forward + backward")
fl = (torch.randn(B, 2, T, H, W, device=dev) * 3).requires_grad_(True)
bk = (torch.randn(B, 2, T, H, W, device=dev) * 3).requires_grad_(True)
mf = torch.rand(B, 1, T, H, W, device=dev).requires_grad_(True)
mb = torch.rand(B, 1, T, H, W, device=dev).requires_grad_(True)
by = 4 * B * T * H * W * (2 * (2 + 2) + 2 + 2 * (2 + 2) + 2)  # fwd reads flows + masks; bwd reads them again, writes 4 + 2 grads


def fc_ours():
    torch.autograd.grad(c2m_b200.flow_consistency_loss(fl, bk, mf, mb), [fl, bk, mf, mb])


def fc_ref():
    torch.autograd.grad(rt.flow_consistency_loss(fl, bk, mf, mb, T), [fl, bk, mf, mb])


a = timed(fc_ours)
print(f"  fused                               {a:8.3f} ms eager   {graphed(fc_ours):8.3f} ms graph   ({by / a / 1e6:6.1f} GB/s of {by / 1e6:.0f} MB)")
print(f"  reference torch composition         {timed(fc_ref, 5):8.3f} ms eager")
# (a process that exits right after a backward can meet PyTorch's autograd worker thread still releasing the last
# graph's tensors while the interpreter finalises: give it a moment, tools/bench_loss_site.py)
torch.cuda.synchronize()
time.sleep(0.2)
