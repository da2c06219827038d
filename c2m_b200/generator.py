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
from .ops import get_corresponding_map, get_grid, get_occlusion_map, grid_sample, resample


def deform_input(inp: torch.Tensor, optical_flow: torch.Tensor) -> torch.Tensor:
    """Same contract as OcclusionAwareGenerator.deform_input (a staticmethod in the reference, generator.py:80-86).
    The reference unpacks the NCHW flow as NHWC and therefore always sends it through F.interpolate(bilinear,
    align_corners=False) without rescaling the values; when the sizes already match that resize is the identity bit
    for bit.  Here the resize happens inside the warp kernel (one launch)."""
    return warp_blend(inp, optical_flow, None, flow_resize="half_pixel")


def apply_optical(self=None, input_ref=None, optical_flow=None, occlusion_map=None):
    """Same contract as OcclusionAwareGenerator.apply_optical (generator.py:88-96): warp `input_ref` by the flow and
    multiply by the occlusion map, both resized to the feature size with bilinear / align_corners=False when their
    sizes differ.  Resize, warp and multiply are ONE kernel; forward saves only the inputs (at their own sizes)."""
    return warp_blend(input_ref, optical_flow, occlusion_map, flow_resize="half_pixel")


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
    materialises T copies of the appearance map folded into the batch (t-major) before warping, resizes the
    motion with `resize_flow` (align_corners=True, values rescaled) and the occlusion with F.interpolate; here the
    kernel reads image n % B, addresses the 5-D motion / occlusion clips through the fold (no torch.cat copies) and
    resizes both on the fly, so the features are read, and their gradient is accumulated, in place.
    sparse_motion [B,2,T,H,W], sparse_occlusion [B,1,T,H,W]."""
    assert sparse_motion.shape[0] == app_features.shape[0] and sparse_motion.shape[2] == num_frames
    # the 5-D clips go to the kernel as they are: frame n = t * B + b reads plane (b, :, t)
    return warp_blend(app_features, sparse_motion, sparse_occlusion, flow_resize="corners_rescaled")


_PATCH_TARGETS = (
    # (module name, attribute) pairs the reference binds by value at import time
    ("utils.ops", "resample"),
    ("utils.ops", "grid_sample"),
    ("utils.ops", "get_grid"),
    ("utils", "resample"),
    ("utils", "grid_sample"),
    ("utils", "get_grid"),
    ("utils.ops", "get_occlusion_map"),
    ("utils.ops", "get_corresponding_map"),
    ("utils", "get_occlusion_map"),
    ("utils", "get_corresponding_map"),
    ("modules.third_party.flow_net.flow_net", "get_occlusion_map"),
    ("modules.generator.generator", "resample"),
    ("modules.motion_estimator.motion_autoencoder", "resample"),
    ("losses.losses", "resample"),
)


def patch_reference(verbose: bool = False):
    """Swap the fused kernels into an already-imported, unmodified C2M source tree.  Returns the
    list of (module, attribute) pairs that were replaced.  The trainer / test.py then run
    unchanged."""
    repl = {"resample": resample, "grid_sample": grid_sample, "get_grid": get_grid,
            "get_occlusion_map": get_occlusion_map, "get_corresponding_map": get_corresponding_map}
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
    # the affine-grid object warp and its objects x T loop (dense_motion.py:94-168)
    dm_mod = sys.modules.get("modules.motion_estimator.dense_motion")
    if dm_mod is not None and hasattr(dm_mod, "DenseMotionNetwork"):
        from . import motion
        cls = dm_mod.DenseMotionNetwork
        cls.warp = staticmethod(motion.affine_warp)
        cls.generate_sparse_motion = motion.generate_sparse_motion
        done.append(("modules.motion_estimator.dense_motion", "DenseMotionNetwork.{warp,generate_sparse_motion}"))
    # the flow-consistency loss (losses.py:115-141): same constructor, fused forward
    loss_mod = sys.modules.get("losses.losses")
    if loss_mod is not None and hasattr(loss_mod, "FlowConsistLoss"):
        from . import loss
        loss_mod.FlowConsistLoss.forward = loss.FlowConsistLoss.forward
        done.append(("losses.losses", "FlowConsistLoss.forward"))
    if verbose:
        for d in done:
            print("c2m_b200: patched %s.%s" % d)
    return done


# ------------------------------------------------------------------------------------------------
# Host module: the reference's occlusion-aware generator with the fused warp swapped in.
#
# Same constructor, same forward signature and -- attribute for attribute -- the same state-dict keys as
# /root/reference/src/modules/generator/generator.py:11-158 (blocks: src/modules/layers/same_block.py:5-23,
# down_block.py:5-24, up_block.py:5-27, residual_block.py:6-31), so reference checkpoints load unchanged
# (SURVEY.md appendix B).  The convolutions / norms stay ordinary PyTorch (cuDNN); only the two warp call
# sites (generator.py:135-137 and :140-145) run the sm_100a kernels.  The SPADE variant (use_spade=True:
# FlowEmbedder + spatially adaptive norms) is not rebuilt here; patch the reference class instead
# (patch_reference) when that variant is needed.
from torch import nn  # noqa: E402


class _ConvNormLeaky(nn.Module):
    """conv -> norm -> LeakyReLU(0.2); `norm` is 'instance' (affine) or 'batch'."""

    def __init__(self, cin, cout, kernel_size, stride, padding, padding_mode, norm):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, kernel_size, stride, padding, padding_mode=padding_mode)
        self.norm = nn.InstanceNorm2d(cout, affine=True) if norm == "instance" else nn.BatchNorm2d(cout)

    def forward(self, x):
        return F.leaky_relu(self.norm(self.conv(x)), 0.2)


class _PreActResidual(nn.Module):
    """BN -> ReLU -> reflect-pad conv, twice, plus the input."""

    def __init__(self, planes, kernel_size, padding):
        super().__init__()
        self.padding = nn.ReflectionPad2d(padding)
        self.conv1 = nn.Conv2d(planes, planes, kernel_size)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size)
        self.norm1 = nn.BatchNorm2d(planes)
        self.norm2 = nn.BatchNorm2d(planes)

    def forward(self, x):
        y = self.conv1(self.padding(F.relu(self.norm1(x))))
        y = self.conv2(self.padding(F.relu(self.norm2(y))))
        return y + x


class _UpsampleConv(nn.Module):
    """x2 bilinear upsample -> conv -> BN -> LeakyReLU(0.2) (parameters live under `main.1` / `main.2`)."""

    def __init__(self, cin, cout, kernel_size, padding, padding_mode):
        super().__init__()
        self.main = nn.Sequential(nn.Upsample(scale_factor=2, mode="bilinear"),
                                  nn.Conv2d(cin, cout, kernel_size, 1, padding, padding_mode=padding_mode),
                                  nn.BatchNorm2d(cout), nn.LeakyReLU(0.2, inplace=True))

    def forward(self, x):
        return self.main(x)


class OcclusionAwareGenerator(nn.Module):
    """Drop-in for the reference generator (non-SPADE variants, `dataset` 'cityscapes' or '...kitti...')."""

    def __init__(self, model_params, flow_params, input_channel, dataset):
        super().__init__()
        if model_params.get("use_spade", False):
            raise NotImplementedError("use_spade=True is not rebuilt in c2m_b200; use patch_reference() on the reference class")
        be, nd = model_params["block_expansion"], model_params["num_down_blocks"]
        cap, pm = model_params["max_expansion"], model_params["padding_mode"]
        self.num_down_blocks, self.dataset, self.flow_params = nd, dataset, flow_params
        width = lambda k: min(cap, be * (2 ** k))  # noqa: E731

        def encoder():
            first = _ConvNormLeaky(input_channel, be, 7, 1, 3, pm, "instance")
            downs = [_ConvNormLeaky(width(k), width(k + 1), 4, 2, 1, pm, "batch") for k in range(nd)]
            return first, downs

        self.first, downs = encoder()
        self.down_blocks = nn.ModuleList(downs)
        if "kitti" in dataset:  # second encoder on the warped frame, merged before decoding
            self.first_warped, downs_w = encoder()
            self.down_blocks_warped = nn.Sequential(*downs_w)
            self.pre_decode = nn.Sequential(_ConvNormLeaky(2 * width(nd), width(nd), 3, 1, 1, pm, "instance"))
        self.up_blocks = nn.ModuleList([_UpsampleConv(width(nd - k), width(nd - k - 1), 3, 1, pm) for k in range(nd)])
        self.middle = nn.Sequential(*[_PreActResidual(width(nd), 3, 1) for _ in range(model_params["num_bottleneck_blocks"])])
        self.final = nn.Sequential(nn.Conv2d(be, 3, kernel_size=7, padding=3), nn.Sigmoid())

    deform_input = staticmethod(deform_input)
    apply_optical = apply_optical

    def forward(self, first_frame, flow, occlusion_map):
        out = self.first(first_frame)
        for blk in self.down_blocks:
            out = blk(out)
        out = self.apply_optical(input_ref=out, optical_flow=flow, occlusion_map=occlusion_map)  # fused warp * mask
        out = self.middle(out)
        if "kitti" in self.dataset:
            warped = self.apply_optical(input_ref=first_frame, optical_flow=flow, occlusion_map=None)
            warped = self.down_blocks_warped(self.first_warped(warped))
            if warped.shape[2:] != occlusion_map.shape[2:]:
                occlusion_map = F.interpolate(occlusion_map, size=warped.shape[2:], mode="bilinear")
            out = self.pre_decode(torch.cat([out, warped * occlusion_map], dim=1))
        for blk in self.up_blocks:
            out = blk(out)
        if out.shape[-2:] != first_frame.shape[-2:]:
            out = F.interpolate(out, list(first_frame.shape[-2:]), mode="bilinear")
        return self.final(out)
