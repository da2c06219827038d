"""Forward-splat occlusion map (reference src/utils/ops.py:205-275): the numpy oracle against golden vectors
produced by the unmodified reference (oracle/make_golden_occmap.py), and -- on the GPU -- the CUDA kernels against
both and against the reference's own torch composition on the device."""
import glob
import os

import numpy as np
import pytest
import torch

import c2m_b200
from oracle import occmap_numpy as on

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "occmap", "*.npz")))
IDS = [os.path.basename(p)[:-4] for p in GOLD]


def _coords(flow):
    b, _, h, w = flow.shape
    jj = np.arange(w, dtype=np.float32).reshape(1, 1, w)
    ii = np.arange(h, dtype=np.float32).reshape(1, h, 1)
    return np.stack([np.broadcast_to(jj, (b, h, w)), np.broadcast_to(ii, (b, h, w))], 1) + flow


def test_fixtures_present():
    assert len(GOLD) >= 7


@pytest.mark.parametrize("path", GOLD, ids=IDS)
def test_numpy_oracle_vs_reference_golden(path):
    d = np.load(path)
    # the reference accumulates with float adds in scatter order, the oracle exactly: a few ulps of the sum
    assert np.abs(on.occlusion_map(d["flow"]) - d["occ"]).max() <= 2e-6
    assert np.abs(on.corresponding_map(_coords(d["flow"])) - d["corr"]).max() <= 2e-6 * max(1.0, d["corr"].max())


def test_oracle_properties():
    # zero flow: every pixel lands on itself with weight 1
    assert np.array_equal(on.occlusion_map(np.zeros((1, 2, 5, 6), np.float32)), np.ones((1, 1, 5, 6), np.float32))
    # mass conservation away from the borders: the four weights of a pixel sum to 1
    f = np.random.default_rng(0).normal(size=(1, 2, 32, 48)).astype(np.float32)
    c = on.corresponding_map(_coords(f))
    inside = (np.abs(f) < 1).all(1)[0][2:-2, 2:-2].sum()  # every pixel whose footprint stays inside contributes 1
    assert c.sum() >= inside - 1e-3


def _torch_reference(flow):
    """The reference composition (ops.py:205-275) written out on the device of `flow` (comparison only)."""
    b, _, h, w = flow.shape
    jj = torch.arange(w, device=flow.device, dtype=torch.float32).view(1, 1, w).expand(b, h, w)
    ii = torch.arange(h, device=flow.device, dtype=torch.float32).view(1, h, 1).expand(b, h, w)
    data = torch.stack([jj, ii], 1) + flow
    x, y = data[:, 0].reshape(b, -1), data[:, 1].reshape(b, -1)
    x1, y1 = torch.floor(x), torch.floor(y)
    xf, yf = x1.clamp(0, w - 1), y1.clamp(0, h - 1)
    x0, y0 = x1 + 1, y1 + 1
    xc, yc = x0.clamp(0, w - 1), y0.clamp(0, h - 1)
    bad = torch.cat([(x0 != xc) | (y0 != yc), (x0 != xc) | (y1 != yf), (x1 != xf) | (y0 != yc), (x1 != xf) | (y1 != yf)], 1)
    idx = torch.cat([xc + yc * w, xc + yf * w, xf + yc * w, xf + yf * w], 1).long()
    val = torch.cat([(1 - (x - xc).abs()) * (1 - (y - yc).abs()), (1 - (x - xc).abs()) * (1 - (y - yf).abs()),
                     (1 - (x - xf).abs()) * (1 - (y - yc).abs()), (1 - (x - xf).abs()) * (1 - (y - yf).abs())], 1)
    val[bad] = 0
    out = torch.zeros(b, h * w, device=flow.device).scatter_add_(1, idx, val)
    return out.view(b, 1, h, w)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=IDS)
def test_cuda_vs_golden(path):
    dev = torch.device("cuda", 0)
    d = np.load(path)
    flow = torch.from_numpy(d["flow"]).to(dev)
    occ = c2m_b200.get_occlusion_map(flow)
    assert occ.shape == (flow.shape[0], 1) + tuple(flow.shape[2:])
    assert np.abs(occ.cpu().numpy() - d["occ"]).max() <= 2e-6
    corr = c2m_b200.get_corresponding_map(torch.from_numpy(_coords(d["flow"])).to(dev))
    assert np.abs(corr.cpu().numpy() - d["corr"]).max() <= 2e-6 * max(1.0, d["corr"].max())
    assert np.abs(corr.cpu().numpy() - on.corresponding_map(_coords(d["flow"]))).max() <= 1e-6 * max(1.0, d["corr"].max())


@pytest.mark.gpu
def test_cuda_full_size_reproducible_and_close_to_the_device_reference():
    dev = torch.device("cuda", 0)
    torch.manual_seed(3)
    flow = torch.randn(10, 2, 256, 512, device=dev) * 6
    flow[0, :, :8] = 1e9        # far out of the image: dropped
    flow[1, 0, 5, 7] = float("nan")
    a = c2m_b200.get_occlusion_map(flow)
    b = c2m_b200.get_occlusion_map(flow)
    assert torch.equal(a, b)  # fixed-point accumulation: bitwise reproducible (scatter_add_ on CUDA is not)
    good = flow.clone()
    good[1, 0, 5, 7] = 1e9  # the reference cannot digest NaN (undefined index cast); ours drops the pixel
    ref = _torch_reference(good).clamp(0, 1)
    assert (a - ref).abs().max().item() <= 2e-6
    assert 0.0 <= a.min().item() and a.max().item() <= 1.0
    with pytest.raises(RuntimeError):
        c2m_b200.get_occlusion_map(flow.cpu())
