"""Stand-in for src/losses/losses.py: the warp call sites (:115-141 FlowConsistLoss, :219-222 the warped term)."""
import torch
import torch.nn.functional as F
from torch import nn

from utils.ops import resample


class FlowConsistLoss(nn.Module):
    def __init__(self, train_params):
        super().__init__()
        self.train_params = train_params

    @staticmethod
    def _flowconsist(flow, flowback, mask_fw=None, mask_bw=None):
        if mask_fw is not None:
            nextloss = (mask_fw * torch.abs(resample(flowback, flow) + flow)).mean()
            prevloss = (mask_bw * torch.abs(resample(flow, flowback) + flowback)).mean()
        else:
            nextloss = torch.abs(resample(flowback, flow) + flow).mean()
            prevloss = torch.abs(resample(flow, flowback) + flowback).mean()
        return prevloss + nextloss

    def forward(self, flow, flowback, mask_fw=None, mask_bw=None):
        fold = lambda t: torch.cat(torch.unbind(t, dim=2), dim=0)  # noqa: E731
        if mask_bw is not None:
            v = self._flowconsist(fold(flow), fold(flowback), fold(mask_fw), fold(mask_bw))
        else:
            v = self._flowconsist(fold(flow), fold(flowback))
        return v * self.train_params["num_predicted_frames"]


def warped_term(source_frame, dense_motion_bw, target_frames):
    T = dense_motion_bw.shape[2]
    warped = torch.cat([resample(source_frame, dense_motion_bw[:, :, i]).unsqueeze(2) for i in range(T)], 2)
    return F.l1_loss(warped, target_frames)
