"""Stand-in for src/modules/generator/generator.py: binds `resample` at import time like the reference (:8) and keeps
the two warp entry points on the class (:80-96)."""
import torch.nn.functional as F
from torch import nn

from utils import resample


class OcclusionAwareGenerator(nn.Module):
    def __init__(self, channels=16):
        super().__init__()
        self.enc = nn.Conv2d(3, channels, 3, stride=8, padding=1)
        self.dec = nn.Conv2d(channels, 3, 3, padding=1)

    @staticmethod
    def deform_input(inp, optical_flow):
        _, h_old, w_old, _ = optical_flow.shape
        _, _, h, w = inp.shape
        if h_old != h or w_old != w:
            optical_flow = F.interpolate(optical_flow, size=(h, w), mode="bilinear")
        return resample(inp, optical_flow)

    def apply_optical(self, input_ref=None, optical_flow=None, occlusion_map=None):
        out = self.deform_input(input_ref, optical_flow)
        if occlusion_map is not None:
            if out.shape[2] != occlusion_map.shape[2] or out.shape[3] != occlusion_map.shape[3]:
                occlusion_map = F.interpolate(occlusion_map, size=out.shape[2:], mode="bilinear")
            out = out * occlusion_map
        return out

    def forward(self, first_frame, flow, occlusion_map):
        out = self.apply_optical(input_ref=self.enc(first_frame), optical_flow=flow, occlusion_map=occlusion_map)
        image = self.apply_optical(input_ref=first_frame, optical_flow=flow)  # the C = 3, no-mask call site (:129-131)
        return F.interpolate(self.dec(out), size=image.shape[2:], mode="bilinear") + image
