"""Fused warped-frame L1 loss (reference src/losses/losses.py:219-222 + L1MaskedLoss losses.py:184-189): the torch
oracle against golden vectors produced by the unmodified reference code (oracle/make_golden_loss.py), and -- on
the GPU -- the CUDA kernels against both and against the reference composition on the device."""
import glob
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import c2m_b200
from oracle import reference_torch as rt

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "loss", "*.npz")))
IDS = [os.path.basename(p)[:-4] for p in GOLD]
LOSS_TOL = 1e-5   # relative, forward
GRAD_TOL = 1e-4   # relative to max |grad|


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    den = b.abs().max().item()
    return (a - b).abs().max().item() / den if den > 0 else (a - b).abs().max().item()


def test_fixtures_present():
    assert len(GOLD) >= 4


@pytest.mark.parametrize("path", GOLD, ids=IDS)
def test_oracle_vs_reference_golden(path):
    d = np.load(path)
    source, targets = torch.from_numpy(d["source"]), torch.from_numpy(d["targets"])
    flows = torch.from_numpy(d["flows"]).requires_grad_(True)
    loss = rt.warped_l1(source, flows, targets)
    (g,) = torch.autograd.grad(loss, [flows])
    assert abs(float(loss) - float(d["loss"])) <= 1e-6 * abs(float(d["loss"]))
    assert rel(g, torch.from_numpy(d["gflows"])) <= 1e-5


def test_cpu_tensors_and_bad_shapes_are_rejected():
    s, f, t = torch.zeros(1, 3, 4, 4), torch.zeros(1, 2, 2, 4, 4), torch.zeros(1, 3, 2, 4, 4)
    with pytest.raises(RuntimeError):
        c2m_b200.warped_l1_loss(s, f, t)  # the reference cannot run this on CPU tensors either (ops.py:189,202)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda", 0)


def _inputs(dev, B, C, T, H, W, seed, amp=3.0):
    g = torch.Generator().manual_seed(seed)
    source = torch.randn(B, C, H, W, generator=g).to(dev)
    targets = torch.randn(B, C, T, H, W, generator=g).to(dev)
    flows = (amp * torch.randn(B, 2, T, H, W, generator=g)).to(dev)
    return source, flows, targets


def _both(source, flows, targets, need_targets=False):
    f1 = flows.detach().clone().requires_grad_(True)
    t1 = targets.detach().clone().requires_grad_(need_targets)
    l1 = c2m_b200.warped_l1_loss(source, f1, t1)
    g1 = torch.autograd.grad(l1, [f1] + ([t1] if need_targets else []))
    f2 = flows.detach().clone().requires_grad_(True)
    t2 = targets.detach().clone().requires_grad_(need_targets)
    l2 = rt.warped_l1(source, f2, t2)
    g2 = torch.autograd.grad(l2, [f2] + ([t2] if need_targets else []))
    return (l1, g1), (l2, g2)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=IDS)
def test_cuda_vs_reference_golden(dev, path):
    d = np.load(path)
    source, targets = torch.from_numpy(d["source"]).to(dev), torch.from_numpy(d["targets"]).to(dev)
    flows = torch.from_numpy(d["flows"]).to(dev).requires_grad_(True)
    loss = c2m_b200.warped_l1_loss(source, flows, targets)
    (g,) = torch.autograd.grad(loss, [flows])
    # CPU goldens: the reference's CPU and CUDA coordinate arithmetic differ -> 1e-4 here
    assert abs(float(loss) - float(d["loss"])) <= 1e-4 * abs(float(d["loss"]))
    assert rel(g, torch.from_numpy(d["gflows"])) <= 2e-4


SHAPES = [(2, 3, 5, 24, 40), (1, 3, 1, 7, 11), (3, 1, 2, 9, 33), (2, 2, 3, 16, 16), (1, 4, 2, 33, 65), (4, 3, 5, 64, 128),
          (1, 3, 2, 2, 2)]


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES, ids=[str(s) for s in SHAPES])
@pytest.mark.parametrize("amp", [0.7, 12.0], ids=["small_flow", "large_flow"])
def test_cuda_vs_device_reference(dev, shape, amp):
    source, flows, targets = _inputs(dev, *shape, seed=sum(shape), amp=amp)
    (l1, g1), (l2, g2) = _both(source, flows, targets, need_targets=True)
    assert abs(float(l1) - float(l2)) <= LOSS_TOL * abs(float(l2))
    assert rel(g1[0], g2[0]) <= GRAD_TOL
    assert rel(g1[1], g2[1]) <= GRAD_TOL


@pytest.mark.gpu
def test_full_size_clip_and_reproducibility(dev):
    """B=8 clips of T=5 frames at 256x512 (BASELINE configs[1] shape of the loss site): against the composition of
    the library's own warp + torch's l1_loss, and bitwise equal run to run (fixed-order double partial sums)."""
    B, C, T, H, W = 8, 3, 5, 256, 512
    source, flows, targets = _inputs(dev, B, C, T, H, W, seed=3, amp=5.0)
    f1 = flows.clone().requires_grad_(True)
    l1 = c2m_b200.warped_l1_loss(source, f1, targets)
    (g1,) = torch.autograd.grad(l1, [f1])
    f2 = flows.clone().requires_grad_(True)
    warped = torch.cat([c2m_b200.resample(source, f2[:, :, t]).unsqueeze(2) for t in range(T)], 2)
    l2 = F.l1_loss(warped, targets)
    (g2,) = torch.autograd.grad(l2, [f2])
    assert abs(float(l1) - float(l2)) <= LOSS_TOL * abs(float(l2))
    assert rel(g1, g2) <= GRAD_TOL
    f3 = flows.clone().requires_grad_(True)
    l3 = c2m_b200.warped_l1_loss(source, f3, targets)
    (g3,) = torch.autograd.grad(l3, [f3])
    assert torch.equal(l1, l3) and torch.equal(g1, g3)


@pytest.mark.gpu
def test_out_of_bounds_flow_and_weighted_loss(dev):
    """Flows far outside the image (border clamp: zero flow gradient where the coordinate is clipped) and an
    upstream factor on the loss (the reference weights this term by 100, yaml `warped`)."""
    source, flows, targets = _inputs(dev, 2, 3, 3, 20, 36, seed=9, amp=60.0)
    f1 = flows.clone().requires_grad_(True)
    (g1,) = torch.autograd.grad(100.0 * c2m_b200.warped_l1_loss(source, f1, targets), [f1])
    f2 = flows.clone().requires_grad_(True)
    (g2,) = torch.autograd.grad(100.0 * rt.warped_l1(source, f2, targets), [f2])
    assert rel(g1, g2) <= GRAD_TOL
    assert (g1 == 0).float().mean() > 0.3  # most coordinates are clipped


@pytest.mark.gpu
def test_source_gradient_route_noncontiguous_and_autocast(dev):
    source, flows, targets = _inputs(dev, 2, 3, 4, 12, 20, seed=5)
    s1, f1 = source.clone().requires_grad_(True), flows.clone().requires_grad_(True)
    l1 = c2m_b200.warped_l1_loss(s1, f1, targets)  # composes the fused warp + l1_loss
    g1 = torch.autograd.grad(l1, [s1, f1])
    s2, f2 = source.clone().requires_grad_(True), flows.clone().requires_grad_(True)
    g2 = torch.autograd.grad(rt.warped_l1(s2, f2, targets), [s2, f2])
    assert rel(g1[0], g2[0]) <= GRAD_TOL and rel(g1[1], g2[1]) <= GRAD_TOL
    # non-contiguous views (the reference slices [:, :, t] out of 5-D tensors)
    big = torch.randn(2, 2, 4, 12, 40, device=dev)
    fv = big[..., ::2]
    assert not fv.is_contiguous()
    la = c2m_b200.warped_l1_loss(source, fv, targets)
    lb = rt.warped_l1(source, fv.contiguous(), targets)
    assert abs(float(la) - float(lb)) <= LOSS_TOL * abs(float(lb))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        lc = c2m_b200.warped_l1_loss(source, flows, targets)
    assert lc.dtype == torch.float32
    assert abs(float(lc) - float(rt.warped_l1(source, flows, targets))) <= LOSS_TOL * abs(float(lc))


@pytest.mark.gpu
def test_empty_clip(dev):
    s, f, t = torch.zeros(0, 3, 4, 4, device=dev), torch.zeros(0, 2, 2, 4, 4, device=dev), torch.zeros(0, 3, 2, 4, 4, device=dev)
    assert torch.isnan(c2m_b200.warped_l1_loss(s, f, t))  # torch: mean of nothing
