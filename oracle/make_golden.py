"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz by running the UNMODIFIED reference sources
from /root/reference on the CPU of the build container.

The reference cannot run on CPU as shipped (``base_grid.cuda(gpu_id)`` with gpu_id == -1,
/root/reference/src/utils/ops.py:189,202), so ``torch.Tensor.cuda`` is replaced by the identity
for the duration of this script -- the reference's own files are imported and executed untouched.
``imageio`` (absent from the image, only used by the visualisation helpers in ops.py) is stubbed.

Run:  python oracle/make_golden.py          (needs /root/reference; the GPU box never runs this)

Each fixture holds the inputs (x, flow, mask, gout) and the reference results: ``out`` from
``utils.ops.resample`` (* mask, generator.py:93) or ``OcclusionAwareGenerator.apply_optical`` and
the autograd gradients gx, gflow, gmask.  These are *CPU* reference results: ATen's CPU path
divides by (size-1)/2 where the CUDA path multiplies by the reciprocal,
so CUDA results are compared against them at 1e-4, and bit-level parity is asserted on the GPU
box against oracle.reference_torch running on the same device.
"""
from __future__ import annotations

import os
import sys
import types
import warnings

import numpy as np
import torch

REF = "/root/reference/src"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _import_reference():
    sys.modules.setdefault("imageio", types.ModuleType("imageio"))
    sys.path.insert(0, REF)
    torch.Tensor.cuda = lambda self, *a, **k: self  # device fix without touching reference files
    import utils.ops as ref_ops  # noqa: E402
    import utils.utils as ref_utils  # noqa: E402
    from modules.generator.generator import OcclusionAwareGenerator  # noqa: E402
    return ref_ops, ref_utils, OcclusionAwareGenerator


def smooth_flow(gen, N, H, W, amp=2.0, noise=0.5):
    ii = torch.arange(H, dtype=torch.float32).view(1, H, 1)
    jj = torch.arange(W, dtype=torch.float32).view(1, 1, W)
    fx = amp * torch.sin(2 * np.pi * ii / max(H / 2.0, 1.0)) * torch.cos(2 * np.pi * jj / max(W / 2.0, 1.0))
    fy = amp * torch.cos(2 * np.pi * ii / max(H / 2.0, 1.0)) * torch.sin(2 * np.pi * jj / max(W / 2.0, 1.0))
    f = torch.stack([fx.expand(N, H, W), fy.expand(N, H, W)], 1)
    return f + noise * torch.randn(N, 2, H, W, generator=gen)


def cases():
    """name -> dict(x, flow, mask|None, kind).  Edge cases listed in SURVEY.md section 8c."""
    g = torch.Generator().manual_seed(20261018)
    rn = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    sig = lambda *s: torch.sigmoid(torch.randn(*s, generator=g))  # noqa: E731
    out = {}
    out["zero_flow_7x11"] = dict(x=rn(2, 3, 7, 11), flow=torch.zeros(2, 2, 7, 11), mask=sig(2, 1, 7, 11))
    fl = torch.zeros(1, 2, 8, 16)
    fl[:, 0] = 2.0
    fl[:, 1] = -1.0
    out["integer_flow_8x16"] = dict(x=rn(1, 4, 8, 16), flow=fl, mask=sig(1, 1, 8, 16))
    fl = torch.zeros(2, 2, 6, 12)
    fl[0] += 0.5
    fl[1] -= 0.5
    out["halfpixel_flow_6x12"] = dict(x=rn(2, 2, 6, 12), flow=fl, mask=None)
    # flows landing exactly on the clip boundaries 0 and W-1 (gradient must vanish there)
    H, W = 5, 9
    jj = torch.arange(W, dtype=torch.float32).view(1, 1, W).expand(1, H, W)
    ii = torch.arange(H, dtype=torch.float32).view(1, H, 1).expand(1, H, W)
    fl = torch.stack([(0.5 * (W - 1) / W) - jj, (H - 1 + 0.5) * (H - 1) / H - ii], 1)
    out["clip_boundary_5x9"] = dict(x=rn(1, 3, H, W), flow=fl.clone(), mask=sig(1, 1, H, W))
    out["large_oob_8x16"] = dict(x=rn(2, 3, 8, 16), flow=rn(2, 2, 8, 16) * 160.0, mask=sig(2, 1, 8, 16))
    fl = smooth_flow(g, 1, 8, 16)
    fl[0, 0, 1, 2] = float("nan")
    fl[0, 1, 3, 4] = float("inf")
    fl[0, 0, 5, 6] = float("-inf")
    out["nonfinite_flow_8x16"] = dict(x=rn(1, 2, 8, 16), flow=fl, mask=sig(1, 1, 8, 16), fwd_only=True)
    out["mask_none_9x13"] = dict(x=rn(2, 5, 9, 13), flow=smooth_flow(g, 2, 9, 13), mask=None)
    out["mask_binary_8x12"] = dict(x=rn(1, 3, 8, 12), flow=smooth_flow(g, 1, 8, 12),
                                   mask=(rn(1, 1, 8, 12) > 0).float())
    out["tiny_2x2"] = dict(x=rn(3, 2, 2, 2), flow=rn(3, 2, 2, 2) * 0.7, mask=sig(3, 1, 2, 2))
    out["w_not_mult4_6x26"] = dict(x=rn(1, 4, 6, 26), flow=smooth_flow(g, 1, 6, 26), mask=sig(1, 1, 6, 26))
    for C in (1, 2, 3, 64):
        out[f"channels_{C}_8x16"] = dict(x=rn(2, C, 8, 16), flow=smooth_flow(g, 2, 8, 16), mask=sig(2, 1, 8, 16))
    out["channels_256_4x8"] = dict(x=rn(1, 256, 4, 8), flow=smooth_flow(g, 1, 4, 8, amp=1.0), mask=sig(1, 1, 4, 8))
    out["channels_512_2x4"] = dict(x=rn(1, 512, 2, 4), flow=rn(1, 2, 2, 4), mask=sig(1, 1, 2, 4))
    # coordinate-rounding sweep over the widths the configs use (pyramid levels of 256/512/832/2048)
    for W in (16, 24, 26, 32, 52, 64, 104, 128, 208, 256, 416, 512, 832, 1024, 2048):
        out[f"width_{W}"] = dict(x=rn(1, 1, 3, W), flow=smooth_flow(g, 1, 3, W, amp=6.0, noise=1.0), mask=None)
    out["cfg1_slice_16x256"] = dict(x=rn(1, 4, 16, 256), flow=smooth_flow(g, 1, 16, 256, amp=8.0, noise=1.0),
                                    mask=sig(1, 1, 16, 256))
    # generator.py:80-96 with a full-resolution flow/mask and a 1/8 feature map (a4/a5 resize path)
    out["apply_optical_resize"] = dict(x=rn(2, 6, 4, 8), flow=smooth_flow(g, 2, 32, 64, amp=3.0),
                                       mask=sig(2, 1, 32, 64), kind="apply_optical")
    out["apply_optical_same"] = dict(x=rn(1, 3, 8, 16), flow=smooth_flow(g, 1, 8, 16), mask=sig(1, 1, 8, 16),
                                     kind="apply_optical")
    # utils.py:346-354 + motion_autoencoder.py:120-125
    out["decoder_scale"] = dict(x=rn(2, 5, 4, 8), flow=smooth_flow(g, 2, 16, 32, amp=4.0), mask=sig(2, 1, 16, 32),
                                kind="decoder")
    return out


def main():
    warnings.simplefilter("ignore")
    ref_ops, ref_utils, Gen = _import_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    gg = torch.Generator().manual_seed(7)
    shim = types.SimpleNamespace(deform_input=Gen.deform_input)
    for name, c in cases().items():
        kind = c.get("kind", "resample")
        x = c["x"].clone().requires_grad_(True)
        flow = c["flow"].clone().requires_grad_(True)
        mask = None if c["mask"] is None else c["mask"].clone().requires_grad_(True)
        if kind == "resample":
            out = ref_ops.resample(x, flow)                       # ops.py:187
            if mask is not None:
                out = out * mask                                  # generator.py:93
        elif kind == "apply_optical":
            out = Gen.apply_optical(shim, x, flow, mask)          # generator.py:88
        elif kind == "decoder":
            import torch.nn.functional as F
            motion = ref_utils.resize_flow(flow, list(x.shape[-2:]))        # utils.py:346
            occ = F.interpolate(mask, size=list(x.shape[-2:]), mode="bilinear")  # motion_autoencoder.py:123
            out = ref_ops.resample(x, motion) * occ               # motion_autoencoder.py:125
        else:
            raise ValueError(kind)
        rec = dict(x=c["x"].numpy(), flow=c["flow"].numpy(), out=out.detach().numpy(), kind=np.array(kind))
        if c["mask"] is not None:
            rec["mask"] = c["mask"].numpy()
        if not c.get("fwd_only", False):
            gout = torch.randn(out.shape, generator=gg)
            ins = [x, flow] + ([mask] if mask is not None else [])
            grads = torch.autograd.grad(out, ins, gout)
            rec.update(gout=gout.numpy(), gx=grads[0].numpy(), gflow=grads[1].numpy())
            if mask is not None:
                rec["gmask"] = grads[2].numpy()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        print(f"{name:28s} out{tuple(out.shape)}")
    # linspace table: the reference's CPU-built base grid for every size any config touches
    sizes = sorted({2, 3, 4, 5, 6, 7, 8, 9, 11, 12, 13, 16, 24, 26, 32, 52, 64, 104, 128, 208, 256, 416, 512, 832,
                    1024, 2048})
    lin = {f"n{n}": ref_ops.get_grid(1, 1, n)[0, 0, 0].numpy() for n in sizes}
    np.savez_compressed(os.path.join(OUT, "base_grid_rows.npz"), **lin)


if __name__ == "__main__":
    main()
