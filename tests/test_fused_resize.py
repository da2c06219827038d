"""GPU parity of the fused flow / mask resize and of the blend operand's fast kernels.

Reference compositions: oracle.reference_torch.apply_optical (generator.py:80-96: F.interpolate bilinear with
align_corners=False, values kept) and .decoder_warp (motion_autoencoder.py:117-125 + utils.py:346-354: align_corners
=True with the values rescaled; occlusion with align_corners=False), run on the same device; gradients by torch
autograd through those compositions (upsample_bilinear2d_backward + the in-place divides).
Tolerances as everywhere: forward <= 1e-5, gradients <= 1e-4, max|a-b| / max|b|.
"""
import pytest
import torch

import c2m_b200
from c2m_b200 import _lib
from oracle import reference_torch as rt

pytestmark = pytest.mark.gpu
FWD_TOL, GRAD_TOL = 1e-5, 1e-4


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda", 0)


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    den = b.abs().max().item()
    num = (a - b).abs().max().item()
    return num / den if den > 0 else num


def leaves(*ts):
    return [None if t is None else t.detach().clone().requires_grad_(True) for t in ts]


CASES = [
    # (N, C, h, w, flow size, mask size)
    (2, 12, 4, 8, (32, 64), (32, 64)),       # 1/8 feature map, the generator's site
    (2, 256, 32, 64, (256, 512), (256, 512)),  # the real cityscapes site (C=256 at 1/8)
    (3, 8, 16, 24, (16, 24), (64, 96)),      # only the mask is resized
    (2, 8, 16, 24, (61, 50), (16, 24)),      # only the flow, non-integer ratio
    (1, 4, 20, 36, (7, 9), (5, 4)),          # up-scaling: many destinations per source pixel
    (2, 16, 9, 13, (40, 50), (23, 31)),      # odd everything
    (2, 4, 2, 3, (5, 7), (3, 2)),            # tiny output
]


@pytest.mark.parametrize("case", CASES, ids=[str(c) for c in CASES])
@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_apply_optical_fused_resize(dev, case, layout):
    N, C, h, w, fs, ms = case
    g = torch.Generator().manual_seed(N * 1000 + C + h)
    x = torch.randn(N, C, h, w, generator=g).to(dev)
    if layout == "nhwc":
        x = x.contiguous(memory_format=torch.channels_last)
    flow = (torch.randn(N, 2, *fs, generator=g) * 3).to(dev)
    mask = torch.rand(N, 1, *ms, generator=g).to(dev)
    gout = torch.randn(N, C, h, w, generator=g).to(dev)
    x1, f1, m1 = leaves(x, flow, mask)
    n0 = _lib.launch_count()
    out = c2m_b200.apply_optical(None, x1, f1, m1)
    # (an NCHW x with C >= 8, C % 4 == 0 is first converted to channels-last: one relayout launch more)
    converted = layout == "nchw" and C >= 8 and C % 4 == 0
    assert _lib.launch_count() - n0 == 1 + converted, "resize + warp + multiply must be ONE forward launch"
    g1 = torch.autograd.grad(out, [x1, f1, m1], gout)
    x2, f2, m2 = leaves(x, flow, mask)
    ref = rt.apply_optical(x2, f2, m2)
    g2 = torch.autograd.grad(ref, [x2, f2, m2], gout)
    assert rel(out, ref) <= FWD_TOL
    for name, a, b in zip(("gx", "gflow", "gmask"), g1, g2):
        assert a.shape == b.shape
        assert rel(a, b) <= GRAD_TOL, f"{name}: {rel(a, b):.3e}"
    # gradient subsets, mask None, deterministic bits
    x3, f3 = leaves(x, flow)
    o3 = c2m_b200.deform_input(x3, f3)
    x4, f4 = leaves(x, flow)
    r3 = rt.deform_input(x4, f4)
    assert rel(o3, r3) <= FWD_TOL
    assert rel(torch.autograd.grad(o3, [f3], gout)[0], torch.autograd.grad(r3, [f4], gout)[0]) <= GRAD_TOL
    d = []
    for _ in range(2):
        xa, fa, ma = leaves(x, flow, mask)
        d.append(torch.autograd.grad(c2m_b200.warp_blend(xa, fa, ma, flow_resize="half_pixel", deterministic=True),
                                     [xa, fa, ma], gout))
    assert all(torch.equal(a, b) for a, b in zip(d[0], d[1]))


DEC = [
    # (B, T, C, h, w, source size)
    (2, 5, 16, 8, 16, (32, 64)),
    (2, 5, 512, 8, 16, (128, 256)),    # the decoder's deepest level
    (1, 5, 64, 64, 128, (128, 256)),   # ... and its shallowest
    (2, 3, 8, 7, 11, (20, 33)),
    (1, 2, 4, 12, 12, (12, 12)),       # same size: the resize is the identity
]


@pytest.mark.parametrize("case", DEC, ids=[str(c) for c in DEC])
@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_decoder_warp_fused_resize(dev, case, layout):
    B, T, C, h, w, ss = case
    g = torch.Generator().manual_seed(B * 100 + C)
    app = torch.randn(B, C, h, w, generator=g).to(dev)
    if layout == "nhwc":
        app = app.contiguous(memory_format=torch.channels_last)
    motion = (torch.randn(B, 2, T, *ss, generator=g) * 4).to(dev)
    occ = torch.rand(B, 1, T, *ss, generator=g).to(dev)
    gout = torch.randn(B * T, C, h, w, generator=g).to(dev)
    a1, m1, o1 = leaves(app, motion, occ)
    out = c2m_b200.decoder_warp(a1, m1, o1, T)
    g1 = torch.autograd.grad(out, [a1, m1, o1], gout)
    a2, m2, o2 = leaves(app, motion, occ)
    ref = rt.decoder_warp(a2, m2, o2, T)
    g2 = torch.autograd.grad(ref, [a2, m2, o2], gout)
    assert rel(out, ref) <= FWD_TOL
    for name, a, b in zip(("gapp", "gmotion", "gocc"), g1, g2):
        assert a.shape == b.shape
        assert rel(a, b) <= GRAD_TOL, f"{name}: {rel(a, b):.3e}"


def test_resize_needs_a_mode(dev):
    x = torch.randn(1, 4, 8, 8, device=dev)
    flow = torch.randn(1, 2, 16, 16, device=dev)
    with pytest.raises(ValueError):
        c2m_b200.warp_blend(x, flow, None)
    with pytest.raises(KeyError):
        c2m_b200.warp_blend(x, flow, None, flow_resize="nearest")


BLEND = [(2, 5, 12, 20), (3, 64, 40, 72), (2, 256, 16, 32), (1, 36, 9, 33), (2, 3, 17, 23)]


@pytest.mark.parametrize("shape", BLEND, ids=[str(s) for s in BLEND])
@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_blend_operand_fast_kernels(dev, shape, layout):
    """out = m*warp + (1-m)*other on the layout-specialised kernels (forward: in the warp kernel; backward: the
    regular kernels + one streaming pass for grad-other and the grad-mask correction) against the oracle, and
    bit-for-bit against the stride-generic kernels' forward."""
    N, C, H, W = shape
    g = torch.Generator().manual_seed(sum(shape))
    fmt = torch.channels_last if layout == "nhwc" else torch.contiguous_format
    x = torch.randn(N, C, H, W, generator=g).to(dev).contiguous(memory_format=fmt)
    other = torch.randn(N, C, H, W, generator=g).to(dev).contiguous(memory_format=fmt)
    flow = (torch.randn(N, 2, H, W, generator=g) * 4).to(dev)
    mask = torch.rand(N, 1, H, W, generator=g).to(dev)
    gout = torch.randn(N, C, H, W, generator=g).to(dev)
    for det in (False, True):
        x1, f1, m1, o1 = leaves(x, flow, mask, other)
        out = c2m_b200.warp_blend(x1, f1, m1, o1, deterministic=det)
        g1 = torch.autograd.grad(out, [x1, f1, m1, o1], gout)
        x2, f2, m2, o2 = leaves(x, flow, mask, other)
        ref = rt.warp_blend(x2, f2, m2, o2)
        g2 = torch.autograd.grad(ref, [x2, f2, m2, o2], gout)
        assert rel(out, ref) <= FWD_TOL
        for name, a, b in zip(("gx", "gflow", "gmask", "gother"), g1, g2):
            assert rel(a, b) <= GRAD_TOL, f"{name}: {rel(a, b):.3e}"
    gen = c2m_b200.warp_blend(x, flow, mask, other, flags=_lib.FLAG_FORCE_GENERIC)
    assert torch.equal(out.detach(), gen)
    # grad-other only / grad-mask only
    x1, f1, m1, o1 = leaves(x, flow, mask, other)
    x1.requires_grad_(False)
    f1.requires_grad_(False)
    out = c2m_b200.warp_blend(x1, f1, m1, o1)
    gm, go = torch.autograd.grad(out, [m1, o1], gout)
    assert rel(gm, g2[2]) <= GRAD_TOL and rel(go, g2[3]) <= GRAD_TOL


@pytest.mark.parametrize("same_size", [False, True])
def test_five_d_clips_equal_the_folded_tensors(dev, same_size):
    """warp_blend on the reference's 5-D clips [B,2,T,h,w] / [B,1,T,h,w] (frame n = t * B + b, no torch.cat copies)
    gives the same bits as on the folded [T*B,...] tensors, and the gradients come back 5-D."""
    B, T, C, H, W = 3, 4, 8, 12, 20
    hs, ws = (H, W) if same_size else (24, 40)
    g = torch.Generator().manual_seed(5 + same_size)
    x = torch.randn(B, C, H, W, generator=g).to(dev)
    motion = (torch.randn(B, 2, T, hs, ws, generator=g) * 3).to(dev)
    occ = torch.rand(B, 1, T, hs, ws, generator=g).to(dev)
    gout = torch.randn(B * T, C, H, W, generator=g).to(dev)
    for mode in ("corners_rescaled", "half_pixel"):
        x1, m1, o1 = leaves(x, motion, occ)
        out5 = c2m_b200.warp_blend(x1, m1, o1, flow_resize=mode)
        g5 = torch.autograd.grad(out5, [x1, m1, o1], gout)
        x2, m2, o2 = leaves(x, motion, occ)
        fold = lambda t: torch.cat(torch.unbind(t, 2), 0)  # noqa: E731
        out4 = c2m_b200.warp_blend(x2, fold(m2), fold(o2), flow_resize=mode)
        g4 = torch.autograd.grad(out4, [x2, m2, o2], gout)
        assert torch.equal(out5, out4)
        assert g5[1].shape == motion.shape and g5[2].shape == occ.shape
        for a, b in zip(g5, g4):
            assert rel(a, b) <= 1e-6
    # no mask, forward only
    assert torch.equal(c2m_b200.warp_blend(x, motion, None, flow_resize="half_pixel"),
                       c2m_b200.warp_blend(x, torch.cat(torch.unbind(motion, 2), 0), None, flow_resize="half_pixel"))
    with pytest.raises(ValueError):
        c2m_b200.warp_blend(x, motion, occ[:, :, 0], flow_resize="half_pixel")
