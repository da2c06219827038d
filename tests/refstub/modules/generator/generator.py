"""Stand-in for the import structure of src/modules/generator/generator.py: the module binds `resample` by value at
import time (the reference does, :8) and the class carries the two warp entry points under the reference's names
(:80-96).  Bodies are this repo's own wording of the oracle's composition; the network around them is a toy."""
import torch.nn.functional as F
from torch import nn

from utils import resample


def _to_size(t, size):
    return t if tuple(t.shape[-2:]) == tuple(size) else F.interpolate(t, size=tuple(size), mode="bilinear")


class OcclusionAwareGenerator(nn.Module):
    def __init__(self, channels=16):
        super().__init__()
        self.enc = nn.Conv2d(3, channels, 3, stride=8, padding=1)
        self.dec = nn.Conv2d(channels, 3, 3, padding=1)

    @staticmethod
    def deform_input(inp, optical_flow):
        return resample(inp, _to_size(optical_flow, inp.shape[-2:]))

    def apply_optical(self, input_ref=None, optical_flow=None, occlusion_map=None):
        warped = self.deform_input(input_ref, optical_flow)
        return warped if occlusion_map is None else warped * _to_size(occlusion_map, warped.shape[-2:])

    def forward(self, first_frame, flow, occlusion_map):
        deep = self.apply_optical(input_ref=self.enc(first_frame), optical_flow=flow, occlusion_map=occlusion_map)
        image = self.apply_optical(input_ref=first_frame, optical_flow=flow)  # the C = 3 call site without a mask
        return F.interpolate(self.dec(deep), size=image.shape[2:], mode="bilinear") + image
