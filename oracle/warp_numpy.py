"""TEST INFRASTRUCTURE -- numpy restatement of the whole warp+blend arithmetic, forward and
backward, including the part that lives in a third-party dependency of the reference.

Where the arithmetic comes from
-------------------------------
* base grid, flow normalisation, grid add: /root/reference/src/utils/ops.py:187-202.
* occlusion multiply: /root/reference/src/modules/generator/generator.py:93.
* bilinear border/zeros sampling and its gradient: PyTorch ATen ``grid_sampler_2d`` /
  ``grid_sampler_2d_backward`` -- NOT under /root/reference (un-vendored dependency; the
  reference pins no version, the installed one is torch 2.11.0+cu128).  The published algorithm
  is restated from ``torch/include/ATen/native/cuda/GridSampler.cuh`` (unnormalize :22-31, clip
  :55-58, clip-with-grad :64-82, safe_downgrade :141-148) and the kernel structure of
  ``aten/src/ATen/native/cuda/GridSampler.cu`` (corner order nw, ne, sw, se; gix/giy
  accumulation; ``gix_mult``).

Two coordinate variants exist because the reference's own CPU and CUDA paths disagree
(SURVEY.md appendix A.3): ``variant='cuda'`` multiplies by the fp32 reciprocal of (size-1)/2
(ATen's CUDA ``div`` by a scalar); ``variant='cpu'`` uses a true division.  Both fuse
``(c+1)*size-1`` into one FMA.  ``dtype=np.float64`` gives the closed-form yardstick
(exact linspace, no fp32 rounding).

Pinned by tests/test_oracle.py against tests/golden/*.npz (generated from the unmodified
reference by oracle/make_golden.py).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def _fma32(a, b, c):
    """fp32 fused multiply-add emulated through float64: the product of two fp32 numbers is
    exact in float64 and, for the magnitudes used here (|a*b|, |c| within a few binades of each
    other), so is the sum, hence a single rounding to fp32 == fmaf."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(F32)


def linspace32(n: int) -> np.ndarray:
    """torch CPU ``linspace(-1, 1, n)`` in float32, bit for bit: step = 2/(n-1) in fp32, first
    half ``fma(step, k, -1)``, second half ``fma(-step, n-1-k, 1)`` (ops.py:198,200 build the
    grid on the CPU even for GPU runs).  n == 1 -> [-1] (ops.py:198)."""
    if n == 1:
        return np.array([-1.0], F32)
    step = F32(F32(2.0) / F32(n - 1))
    k = np.arange(n, dtype=np.int64)
    lo = _fma32(step, k.astype(F32), F32(-1.0))
    hi = _fma32(-step, (n - 1 - k).astype(F32), F32(1.0))
    return np.where(k < n // 2, lo, hi).astype(F32)


def source_coords(flow: np.ndarray, H: int, W: int, variant: str = "cuda", dtype=F32):
    """Unclipped source coordinates (ix, iy), each [N,H,W], for pixel flow [N,2,H,W]
    (channel 0 = x).  ops.py:190-191 + ATen unnormalize (align_corners=False)."""
    fx, fy = flow[:, 0], flow[:, 1]
    if dtype == np.float64:
        gx = (np.linspace(-1.0, 1.0, W) if W > 1 else np.array([-1.0]))[None, None, :]
        gy = (np.linspace(-1.0, 1.0, H) if H > 1 else np.array([-1.0]))[None, :, None]
        with np.errstate(divide="ignore", invalid="ignore"):
            cx = gx + fx.astype(np.float64) / ((W - 1.0) / 2.0)
            cy = gy + fy.astype(np.float64) / ((H - 1.0) / 2.0)
        return ((cx + 1.0) * W - 1.0) / 2.0, ((cy + 1.0) * H - 1.0) / 2.0
    gx = linspace32(W)[None, None, :]
    gy = linspace32(H)[None, :, None]
    bw, bh = F32((W - 1.0) / 2.0), F32((H - 1.0) / 2.0)
    fx = fx.astype(F32)
    fy = fy.astype(F32)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        if variant == "cuda":
            nx = (fx * (F32(1.0) / bw)).astype(F32)
            ny = (fy * (F32(1.0) / bh)).astype(F32)
        elif variant == "cpu":
            nx = (fx / bw).astype(F32)
            ny = (fy / bh).astype(F32)
        else:
            raise ValueError(variant)
        cx = (gx + nx).astype(F32)
        cy = (gy + ny).astype(F32)
        # both ATen builds contract (c+1)*size-1 into one FMA (probed: the unfused form is
        # 2e-5 away from torch-CPU at W=832, the fused one 1e-7)
        ix = (_fma32((cx + F32(1.0)).astype(F32), F32(W), F32(-1.0)) * F32(0.5)).astype(F32)
        iy = (_fma32((cy + F32(1.0)).astype(F32), F32(H), F32(-1.0)) * F32(0.5)).astype(F32)
    return ix, iy


def _clip(c, size, padding, dtype):
    """Returns (clipped coordinate, d clipped / d unclipped).  Border: GridSampler.cuh:55-82
    (NaN -> 0 through max(); zero gradient when c <= 0 or c >= size-1).  Zeros: no clip;
    non-finite / huge -> -100 (:141-148)."""
    c = c.astype(dtype)
    if padding == "border":
        with np.errstate(invalid="ignore"):
            inside = (c > 0) & (c < size - 1)
            cc = np.where(np.isnan(c), dtype(0), np.minimum(dtype(size - 1), np.maximum(c, dtype(0))))
        g = inside.astype(dtype)
    elif padding == "zeros":
        with np.errstate(invalid="ignore"):
            bad = ~np.isfinite(c) | (c > 2147483646.0) | (c < -2147483648.0)
        cc = np.where(bad, dtype(-100.0), c)
        g = np.ones_like(cc)
    else:
        raise ValueError(padding)
    return cc.astype(dtype), g


def _corners(ix, iy, H, W):
    x0 = np.floor(ix)
    y0 = np.floor(iy)
    x1 = x0 + 1
    y1 = y0 + 1
    w_nw = (x1 - ix) * (y1 - iy)
    w_ne = (ix - x0) * (y1 - iy)
    w_sw = (x1 - ix) * (iy - y0)
    w_se = (ix - x0) * (iy - y0)
    xi0, yi0, xi1, yi1 = (a.astype(np.int64) for a in (x0, y0, x1, y1))
    corners = []
    for (yy, xx, ww) in ((yi0, xi0, w_nw), (yi0, xi1, w_ne), (yi1, xi0, w_sw), (yi1, xi1, w_se)):
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        corners.append((np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1), ww, ok))
    return corners, (x0, y0, x1, y1)


def warp_blend_forward(x, flow, mask=None, other=None, padding="border", variant="cuda", dtype=F32):
    """out[n,c,i,j] = mask * bilinear(x[n,c], ix, iy) (+ (1-mask)*other).  x [N,C,H,W], flow
    [N,2,H,W], mask [N,1,H,W] or None.  Returns (out, warped)."""
    N, C, H, W = x.shape
    ix, iy = source_coords(flow, H, W, variant, dtype)
    ix, _ = _clip(ix, W, padding, dtype)
    iy, _ = _clip(iy, H, padding, dtype)
    corners, _ = _corners(ix, iy, H, W)
    n_idx = np.arange(N)[:, None, None]
    xs = x.astype(dtype)
    warped = np.zeros((N, C, H, W), dtype)
    for (yy, xx, ww, ok) in corners:
        v = xs[n_idx, :, yy, xx]  # [N,H,W,C]
        v = np.where(ok[..., None], v, dtype(0))
        warped += np.moveaxis((v * ww[..., None].astype(dtype)).astype(dtype), -1, 1)
    warped = warped.astype(dtype)
    if mask is None:
        return warped, warped
    m = mask.astype(dtype)
    if other is None:
        return (warped * m).astype(dtype), warped
    return ((warped * m).astype(dtype) + ((dtype(1) - m) * other.astype(dtype)).astype(dtype)).astype(dtype), warped


def warp_blend_backward(x, flow, mask, gout, other=None, padding="border", variant="cuda", dtype=F32):
    """Gradients of warp_blend_forward w.r.t. x, flow, mask (and other).  Reductions are carried
    in float64 (the oracle is the yardstick, not a bit-replica of atomics order).
    Returns dict(gx, gflow, gmask, gother)."""
    N, C, H, W = x.shape
    ixu, iyu = source_coords(flow, H, W, variant, dtype)
    ix, cgx = _clip(ixu, W, padding, dtype)
    iy, cgy = _clip(iyu, H, padding, dtype)
    corners, (x0, y0, x1, y1) = _corners(ix, iy, H, W)
    xs = x.astype(np.float64)
    go = gout.astype(np.float64)
    m = None if mask is None else mask.astype(np.float64)
    g = go if m is None else go * m  # [N,C,H,W], mul backward (generator.py:93)
    n_idx = np.arange(N)[:, None, None]
    vals = []
    gx = np.zeros(N * C * H * W, np.float64)
    cbase = (np.arange(N)[:, None, None, None] * C + np.arange(C)[None, :, None, None]) * (H * W)
    for (yy, xx, ww, ok) in corners:
        v = np.moveaxis(np.where(ok[..., None], xs[n_idx, :, yy, xx], 0.0), -1, 1)  # [N,C,H,W]
        vals.append(v)
        contrib = g * (ww.astype(np.float64) * ok)[:, None]
        dest = cbase + (yy * W + xx)[:, None]
        gx += np.bincount(dest.ravel(), weights=contrib.ravel(), minlength=gx.size)
    gx = gx.reshape(N, C, H, W)
    v_nw, v_ne, v_sw, v_se = vals
    ixd, iyd = ix.astype(np.float64)[:, None], iy.astype(np.float64)[:, None]
    x0d, y0d, x1d, y1d = (a.astype(np.float64)[:, None] for a in (x0, y0, x1, y1))
    # ATen grid_sampler_2d_backward: d out / d ix, d out / d iy summed over channels
    gix = (g * ((v_ne - v_nw) * (y1d - iyd) + (v_se - v_sw) * (iyd - y0d))).sum(1)
    giy = (g * ((v_sw - v_nw) * (x1d - ixd) + (v_se - v_ne) * (ixd - x0d))).sum(1)
    with np.errstate(divide="ignore", invalid="ignore"):
        # gix_mult = size/2 * clip-grad; ops.py:190 contributes 1/((size-1)/2)
        sx = (W / 2.0) / ((W - 1.0) / 2.0) if W > 1 else np.inf
        sy = (H / 2.0) / ((H - 1.0) / 2.0) if H > 1 else np.inf
        gflow = np.stack([gix * cgx.astype(np.float64) * sx, giy * cgy.astype(np.float64) * sy], 1)
    res = {"gx": gx, "gflow": gflow, "gmask": None, "gother": None}
    if m is not None:
        w00, w01, w10, w11 = (c[2].astype(np.float64)[:, None] for c in corners)
        warped = v_nw * w00 + v_ne * w01 + v_sw * w10 + v_se * w11
        if other is None:
            res["gmask"] = (go * warped).sum(1, keepdims=True)
        else:
            res["gmask"] = (go * (warped - other.astype(np.float64))).sum(1, keepdims=True)
            res["gother"] = go * (1.0 - m)
    elif other is not None:
        raise ValueError("other requires mask")
    return res


def rel_err(a, b) -> float:
    """The repo-wide error definition: max|a-b| / max|b|."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    denom = float(np.max(np.abs(b))) if b.size else 0.0
    if denom == 0.0:
        return float(np.max(np.abs(a - b))) if a.size else 0.0
    return float(np.max(np.abs(a - b)) / denom)
