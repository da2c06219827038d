"""TEST INFRASTRUCTURE -- numpy restatement of the reference's forward-splat occlusion map
(/root/reference/src/utils/ops.py:205-275: get_corresponding_map, mesh_grid, get_occlusion_map).

Every pixel (i, j) of a flow field is moved to (j + fx, i + fy) -- plain pixel coordinates, none of the
align_corners quirks of the warp path -- and its bilinear weights are scatter-added onto the four surrounding
pixels; corners that fall outside the image are dropped (ops.py:221-231,247); the occlusion map is that sum
clamped to [0, 1] (ops.py:275).  Pinned by tests/golden/occmap/*.npz, generated from the unmodified reference by
oracle/make_golden_occmap.py.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def corresponding_map(data: np.ndarray) -> np.ndarray:
    """ops.py:205-251.  data [B,2,H,W] absolute (unnormalised) coordinates -> [B,1,H,W]."""
    data = np.asarray(data, dtype=F32)
    b, _, h, w = data.shape
    x = data[:, 0].reshape(b, -1)
    y = data[:, 1].reshape(b, -1)
    x1 = np.floor(x)
    xf = np.clip(x1, 0, w - 1)
    y1 = np.floor(y)
    yf = np.clip(y1, 0, h - 1)
    x0 = x1 + F32(1)
    xc = np.clip(x0, 0, w - 1)
    y0 = y1 + F32(1)
    yc = np.clip(y0, 0, h - 1)
    xc_out, yc_out, xf_out, yf_out = x0 != xc, y0 != yc, x1 != xf, y1 != yf
    out = np.zeros((b, h * w), dtype=np.float64)  # exact accumulation: the reference's atomics have no fixed order
    one = F32(1)
    for (xa, ya, bad) in ((xc, yc, xc_out | yc_out), (xc, yf, xc_out | yf_out), (xf, yc, xf_out | yc_out),
                          (xf, yf, xf_out | yf_out)):
        val = ((one - np.abs(x - xa)) * (one - np.abs(y - ya))).astype(F32)
        val[bad] = 0
        idx = (xa + ya * F32(w)).astype(np.int64)
        for k in range(b):
            np.add.at(out[k], idx[k], val[k].astype(np.float64))
    return out.astype(F32).reshape(b, 1, h, w)


def occlusion_map(flow: np.ndarray) -> np.ndarray:
    """ops.py:263-275: corresponding map of (mesh grid + flow), clamped to [0, 1]."""
    flow = np.asarray(flow, dtype=F32)
    b, _, h, w = flow.shape
    jj = np.arange(w, dtype=F32).reshape(1, 1, w)
    ii = np.arange(h, dtype=F32).reshape(1, h, 1)
    base = np.stack([np.broadcast_to(jj, (b, h, w)), np.broadcast_to(ii, (b, h, w))], 1)
    return np.clip(corresponding_map(base + flow), 0.0, 1.0)
