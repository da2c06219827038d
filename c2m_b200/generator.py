"""Host-side mirror of the reference generator's warp entry points
(/root/reference/src/modules/generator/generator.py:80-96) plus the hooks that swap the fused
kernels into an unmodified C2M checkout.

    deform_input(inp, optical_flow)                                   generator.py:80-86
    apply_optical(self, input_ref=None, optical_flow=None,
                  occlusion_map=None)                                 generator.py:88-96
    decoder_warp(...)                                                 motion_autoencoder.py:117-125
    patch_reference()                                                 monkey-patches the names the
                                                                      reference binds at import time
"""
from __future__ import annotations

import sys

import torch
import torch.nn.functional as F

from .functional import warp_blend
from .ops import get_grid, grid_sample, resample


def _flow_for(inp: torch.Tensor, optical_flow: torch.Tensor) -> torch.Tensor:
    """generator.py:82-85.  The reference unpacks the NCHW flow as NHWC and therefore always sends
    it through F.interpolate(bilinear, align_corners=False) without rescaling the values; when the
    sizes already match that resize is the identity bit for bit, so
    it is skipped here."""
    h, w = inp.shape[2:]
    if optical_flow.shape[2] != h or optical_flow.shape[3] != w:
        optical_flow = F.interpolate(optical_flow, size=(h, w), mode="bilinear")
    return optical_flow


def deform_input(inp: torch.Tensor, optical_flow: torch.Tensor) -> torch.Tensor:
    """Same contract as OcclusionAwareGenerator.deform_input (a staticmethod in the reference)."""
    return resample(inp, _flow_for(inp, optical_flow))


def apply_optical(self=None, input_ref=None, optical_flow=None, occlusion_map=None):
    """Same contract as OcclusionAwareGenerator.apply_optical: warp `input_ref` by the flow and
    multiply by the occlusion map (resized with bilinear/align_corners=False when its size
    differs).  The warp and the multiply are one kernel; forward saves only the inputs."""
    flow = _flow_for(input_ref, optical_flow)
    if occlusion_map is not None:
        if input_ref.shape[2] != occlusion_map.shape[2] or input_ref.shape[3] != occlusion_map.shape[3]:
            occlusion_map = F.interpolate(occlusion_map, size=input_ref.shape[2:], mode="bilinear")
    return warp_blend(input_ref, flow, occlusion_map)


def resize_flow(flow: torch.Tensor, new_shape) -> torch.Tensor:
    """utils.py:346-354 (align_corners=True bilinear resize, values rescaled by new/old)."""
    _, _, h, w = flow.shape
    new_h, new_w = new_shape
    out = F.interpolate(flow, (new_h, new_w), mode="bilinear", align_corners=True)
    out[:, 0] /= w / float(new_w)
    out[:, 1] /= h / float(new_h)
    return out


def decoder_warp(app_features: torch.Tensor, sparse_motion: torch.Tensor, sparse_occlusion: torch.Tensor,
                 num_frames: int) -> torch.Tensor:
    """One scale of DenseMotionDecoder.forward (motion_autoencoder.py:117-125).  The reference
    materialises T copies of the appearance map folded into the batch (t-major) before warping;
    here the kernel reads image n % B instead, so the features are read, and their gradient is
    accumulated, in place.  sparse_motion [B,2,T,H,W], sparse_occlusion [B,1,T,H,W]."""
    nh, nw = app_features.shape[-2:]
    motion = resize_flow(torch.cat(torch.unbind(sparse_motion, 2), 0), [nh, nw])
    occ = F.interpolate(torch.cat(torch.unbind(sparse_occlusion, 2), 0), size=[nh, nw], mode="bilinear")
    assert motion.shape[0] == app_features.shape[0] * num_frames
    return warp_blend(app_features, motion, occ)


_PATCH_TARGETS = (
    # (module name, attribute) pairs the reference binds by value at import time
    ("utils.ops", "resample"),
    ("utils.ops", "grid_sample"),
    ("utils.ops", "get_grid"),
    ("utils", "resample"),
    ("utils", "grid_sample"),
    ("utils", "get_grid"),
    ("modules.generator.generator", "resample"),
    ("modules.motion_estimator.motion_autoencoder", "resample"),
    ("losses.losses", "resample"),
)


def patch_reference(verbose: bool = False):
    """Swap the fused kernels into an already-imported, unmodified C2M source tree.  Returns the
    list of (module, attribute) pairs that were replaced.  The trainer / test.py then run
    unchanged."""
    repl = {"resample": resample, "grid_sample": grid_sample, "get_grid": get_grid}
    done = []
    for mod_name, attr in _PATCH_TARGETS:
        mod = sys.modules.get(mod_name)
        if mod is not None and hasattr(mod, attr):
            setattr(mod, attr, repl[attr])
            done.append((mod_name, attr))
    gen_mod = sys.modules.get("modules.generator.generator")
    if gen_mod is not None and hasattr(gen_mod, "OcclusionAwareGenerator"):
        cls = gen_mod.OcclusionAwareGenerator
        cls.deform_input = staticmethod(deform_input)
        cls.apply_optical = apply_optical
        done.append(("modules.generator.generator", "OcclusionAwareGenerator.{deform_input,apply_optical}"))
    if verbose:
        for d in done:
            print("c2m_b200: patched %s.%s" % d)
    return done
