"""CPU tests: the oracle (oracle/) against the golden vectors produced from the unmodified
reference (oracle/make_golden.py), plus internal consistency of the two restatements."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import reference_torch as rt
from oracle import warp_numpy as wn

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
CASES = [g for g in GOLD if not g.endswith("base_grid_rows.npz")]


def _load(path):
    d = np.load(path)
    return {k: d[k] for k in d.files}


def _ids(paths):
    return [os.path.basename(p)[:-4] for p in paths]


def test_golden_present():
    assert len(CASES) >= 30 and any(g.endswith("base_grid_rows.npz") for g in GOLD)


def test_linspace_matches_reference_base_grid(golden_dir):
    d = np.load(os.path.join(golden_dir, "base_grid_rows.npz"))
    for k in d.files:
        assert np.array_equal(d[k], wn.linspace32(int(k[1:]))), k


@pytest.mark.parametrize("n", list(range(1, 300)) + [416, 511, 512, 513, 832, 1024, 2047, 2048, 4096])
def test_linspace_matches_torch_bitwise(n):
    ref = torch.linspace(-1, 1, n).numpy() if n > 1 else np.array([-1.0], np.float32)
    assert np.array_equal(ref, wn.linspace32(n))


@pytest.mark.parametrize("path", CASES, ids=_ids(CASES))
def test_numpy_oracle_vs_golden(path):
    d = _load(path)
    if str(d["kind"]) != "resample":
        pytest.skip("composite case: covered by the torch restatement")
    mask = d.get("mask")
    out, _ = wn.warp_blend_forward(d["x"], d["flow"], mask, variant="cpu")
    assert wn.rel_err(out, d["out"]) <= 1e-6
    # the CUDA-arithmetic variant differs from the CPU reference only by the reciprocal multiply
    out_c, _ = wn.warp_blend_forward(d["x"], d["flow"], mask, variant="cuda")
    assert wn.rel_err(out_c, d["out"]) <= 1e-4
    if "gout" in d:
        r = wn.warp_blend_backward(d["x"], d["flow"], mask, d["gout"], variant="cpu")
        assert wn.rel_err(r["gx"], d["gx"]) <= 2e-6
        assert wn.rel_err(r["gflow"], d["gflow"]) <= 2e-6
        if mask is not None:
            assert wn.rel_err(r["gmask"], d["gmask"]) <= 2e-6


@pytest.mark.parametrize("path", CASES, ids=_ids(CASES))
def test_torch_restatement_vs_golden(path):
    d = _load(path)
    kind = str(d["kind"])
    x = torch.from_numpy(d["x"]).requires_grad_(True)
    flow = torch.from_numpy(d["flow"]).requires_grad_(True)
    mask = torch.from_numpy(d["mask"]).requires_grad_(True) if "mask" in d else None
    if kind == "resample":
        out = rt.warp_blend(x, flow, mask)
    elif kind == "apply_optical":
        out = rt.apply_optical(x, flow, mask)
    else:
        import torch.nn.functional as F
        out = rt.resample(x, rt.resize_flow(flow, list(x.shape[-2:]))) * F.interpolate(
            mask, size=list(x.shape[-2:]), mode="bilinear")
    # same code path as the reference; not asserted bit-equal because ATen's CPU kernels pick
    # their vector ISA (and with it FMA contraction) per host
    assert wn.rel_err(out.detach().numpy(), d["out"]) <= 1e-6
    if "gout" in d:
        ins = [x, flow] + ([mask] if mask is not None else [])
        g = torch.autograd.grad(out, ins, torch.from_numpy(d["gout"]))
        assert wn.rel_err(g[0].numpy(), d["gx"]) <= 1e-6
        assert wn.rel_err(g[1].numpy(), d["gflow"]) <= 1e-6
        if mask is not None:
            assert wn.rel_err(g[2].numpy(), d["gmask"]) <= 1e-6


def test_zero_flow_is_not_identity():
    # convention mismatch, ix = j*W/(W-1) - 0.5
    x = np.random.default_rng(0).standard_normal((1, 2, 8, 16)).astype(np.float32)
    out, _ = wn.warp_blend_forward(x, np.zeros((1, 2, 8, 16), np.float32))
    assert np.abs(out - x).max() > 0.1
    ix, _ = wn.source_coords(np.zeros((1, 2, 8, 16), np.float32), 8, 16)
    j = np.arange(16)
    assert np.allclose(ix[0, 0], j * 16 / 15.0 - 0.5, atol=1e-5)


def test_fp64_closed_form_close_to_fp32():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2, 3, 16, 32)).astype(np.float32)
    flow = (rng.standard_normal((2, 2, 16, 32)) * 3).astype(np.float32)
    mask = rng.random((2, 1, 16, 32)).astype(np.float32)
    o32, _ = wn.warp_blend_forward(x, flow, mask)
    o64, _ = wn.warp_blend_forward(x, flow, mask, dtype=np.float64)
    assert wn.rel_err(o32, o64) < 1e-5


def test_numpy_backward_matches_autograd_fp64():
    """The numpy backward (restated ATen grid_sampler_2d_backward) against autograd of the torch
    restatement, border and zeros padding, with and without `other`."""
    import torch.nn.functional as F
    rng = np.random.default_rng(2)
    N, C, H, W = 2, 3, 9, 14
    x = rng.standard_normal((N, C, H, W)).astype(np.float32)
    flow = (rng.standard_normal((N, 2, H, W)) * 4).astype(np.float32)
    mask = rng.random((N, 1, H, W)).astype(np.float32)
    other = rng.standard_normal((N, C, H, W)).astype(np.float32)
    gout = rng.standard_normal((N, C, H, W)).astype(np.float32)
    for padding in ("border", "zeros"):
        for use_other in (False, True):
            tx = torch.from_numpy(x).requires_grad_(True)
            tf = torch.from_numpy(flow).requires_grad_(True)
            tm = torch.from_numpy(mask).requires_grad_(True)
            to = torch.from_numpy(other).requires_grad_(True)
            grid = rt.base_grid(N, H, W, "cpu")
            nf = torch.cat([tf[:, 0:1] / ((W - 1.0) / 2.0), tf[:, 1:2] / ((H - 1.0) / 2.0)], 1)
            warped = F.grid_sample(tx, (grid + nf).permute(0, 2, 3, 1), mode="bilinear", padding_mode=padding,
                                   align_corners=False)
            out = warped * tm + ((1 - tm) * to if use_other else 0)
            gs = torch.autograd.grad(out, [tx, tf, tm] + ([to] if use_other else []), torch.from_numpy(gout))
            o, _ = wn.warp_blend_forward(x, flow, mask, other if use_other else None, padding=padding, variant="cpu")
            assert wn.rel_err(o, out.detach().numpy()) < 1e-6
            r = wn.warp_blend_backward(x, flow, mask, gout, other if use_other else None, padding=padding,
                                       variant="cpu")
            assert wn.rel_err(r["gx"], gs[0].numpy()) < 2e-6
            assert wn.rel_err(r["gflow"], gs[1].numpy()) < 2e-6
            assert wn.rel_err(r["gmask"], gs[2].numpy()) < 2e-6
            if use_other:
                assert wn.rel_err(r["gother"], gs[3].numpy()) < 2e-6


def test_linearity_and_mask_scaling_properties():
    rng = np.random.default_rng(3)
    x1 = rng.standard_normal((1, 2, 6, 10)).astype(np.float64)
    x2 = rng.standard_normal((1, 2, 6, 10)).astype(np.float64)
    flow = rng.standard_normal((1, 2, 6, 10)).astype(np.float32) * 2
    mask = rng.random((1, 1, 6, 10))
    f = lambda x, m: wn.warp_blend_forward(x, flow, m, dtype=np.float64)[0]  # noqa: E731
    assert np.allclose(f(2 * x1 + 3 * x2, mask), 2 * f(x1, mask) + 3 * f(x2, mask), atol=1e-12)
    assert np.allclose(f(x1, 0.5 * mask), 0.5 * f(x1, mask), atol=1e-12)


def test_integer_shift_away_from_borders():
    """A flow that makes ix an exact integer shift reproduces shifted pixels: with ix = (j+fx)*W/(W-1)-0.5
    choose fx so that ix = j + 2."""
    H, W = 6, 16
    rng = np.random.default_rng(4)
    x = rng.standard_normal((1, 1, H, W))
    j = np.arange(W)[None, None, :]
    i = np.arange(H)[None, :, None]
    fx = (j + 2 + 0.5) * (W - 1) / W - j
    fy = (i + 0.5) * (H - 1) / H - i
    flow = np.stack([np.broadcast_to(fx, (1, H, W)), np.broadcast_to(fy, (1, H, W))], 1)
    out, _ = wn.warp_blend_forward(x, flow, dtype=np.float64)
    assert np.allclose(out[..., : W - 2], x[..., 2:], atol=1e-9)


# ------------------------------------------------------------------------------------------------
# hypothesis-driven properties of the oracle: random shapes, flows and seeds
from hypothesis import given, settings, strategies as st  # noqa: E402

_shape = st.tuples(st.integers(1, 3), st.integers(1, 5), st.integers(2, 12), st.integers(2, 20))


@settings(max_examples=30, deadline=None)
@given(shape=_shape, seed=st.integers(0, 2 ** 16), amp=st.sampled_from([0.3, 2.0, 9.0, 60.0]),
       padding=st.sampled_from(["border", "zeros"]))
def test_property_linearity_mask_scaling_and_adjointness(shape, seed, amp, padding):
    """For any shape / flow: the forward is linear in x and in the mask, and grad-input is the adjoint of the
    forward (<gout, F(x)> == <F^T(gout), x>), in fp64 where the identities hold to rounding."""
    N, C, H, W = shape
    rng = np.random.default_rng(seed)
    x1, x2 = rng.standard_normal((2, N, C, H, W))
    flow = (rng.standard_normal((N, 2, H, W)) * amp).astype(np.float32)
    mask = rng.random((N, 1, H, W))
    gout = rng.standard_normal((N, C, H, W))
    f = lambda x, m: wn.warp_blend_forward(x, flow, m, padding=padding, dtype=np.float64)[0]  # noqa: E731
    assert np.allclose(f(2 * x1 - 3 * x2, mask), 2 * f(x1, mask) - 3 * f(x2, mask), atol=1e-10)
    assert np.allclose(f(x1, 0.25 * mask), 0.25 * f(x1, mask), atol=1e-12)
    r = wn.warp_blend_backward(x1, flow, mask, gout, padding=padding, dtype=np.float64)
    lhs, rhs = float((gout * f(x1, mask)).sum()), float((r["gx"] * x1).sum())
    assert abs(lhs - rhs) <= 1e-9 * (np.abs(gout).sum() + 1.0)
    # grad-mask is the plain warp weighted by gout (mul backward, generator.py:93)
    assert np.allclose(r["gmask"], (gout * f(x1, np.ones_like(mask))).sum(1, keepdims=True), atol=1e-10)


@settings(max_examples=20, deadline=None)
@given(shape=_shape, seed=st.integers(0, 2 ** 16))
def test_property_border_clamp_is_idempotent_for_far_flows(shape, seed):
    """Border padding: once a sample is outside the image, pushing it further out changes nothing, and the flow
    gradient there is zero (clip_coordinates_set_grad)."""
    N, C, H, W = shape
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((N, C, H, W)).astype(np.float32)
    sign = np.where(rng.random((N, 2, H, W)) < 0.5, -1.0, 1.0)
    far = (sign * 4.0 * max(H, W)).astype(np.float32)
    o1, _ = wn.warp_blend_forward(x, far)
    o2, _ = wn.warp_blend_forward(x, (3 * far).astype(np.float32))
    assert np.array_equal(o1, o2)
    r = wn.warp_blend_backward(x, far, None, np.ones_like(x))
    assert np.all(r["gflow"] == 0)
