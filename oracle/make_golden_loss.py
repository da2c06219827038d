"""TEST INFRASTRUCTURE -- generates tests/golden/loss/*.npz by running the UNMODIFIED reference code of the
`warped` training-loss term on the CPU of the build container: ``losses.losses.resample`` (the name
/root/reference/src/losses/losses.py:6 binds) called once per predicted frame, ``torch.cat`` on the frame axis and
``losses.losses.L1MaskedLoss`` (losses.py:180-189) -- the statements of losses.py:219-222 with the reference's own
callables.  As in oracle/make_golden.py, ``torch.Tensor.cuda`` is the identity for the duration of the script
(ops.py:189,202 hard-code ``.cuda(gpu_id)``) and ``imageio`` is stubbed; no reference file is modified.

Run:  python oracle/make_golden_loss.py      (needs /root/reference; the GPU box never runs this)

Each fixture: source [B,C,H,W], flows [B,2,T,H,W], targets [B,C,T,H,W], the loss, and d loss / d flows from
autograd through the reference code.  CPU results (true division by (size-1)/2): CUDA
results are compared at 1e-4, bit-level parity is asserted on the GPU against oracle.reference_torch there.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/src"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "loss")


def main():
    sys.modules.setdefault("imageio", types.ModuleType("imageio"))
    sys.path.insert(0, REF)
    torch.Tensor.cuda = lambda self, *a, **k: self
    import losses.losses as ref_losses
    os.makedirs(OUT, exist_ok=True)
    l1 = ref_losses.L1MaskedLoss()
    g = torch.Generator().manual_seed(20261020)
    cases = {"rgb_t3_12x20": (2, 3, 3, 12, 20, 2.0), "rgb_t5_16x32": (1, 3, 5, 16, 32, 4.0),
             "two_channels_t2_7x11": (2, 2, 2, 7, 11, 1.5), "one_frame_9x13": (3, 3, 1, 9, 13, 6.0)}
    for name, (B, C, T, H, W, amp) in cases.items():
        source = torch.randn(B, C, H, W, generator=g)
        targets = torch.randn(B, C, T, H, W, generator=g)
        flows = (amp * torch.randn(B, 2, T, H, W, generator=g)).requires_grad_(True)
        warped_frames = torch.cat([torch.unsqueeze(ref_losses.resample(source, flows[:, :, i, ...]), 2)
                                   for i in range(T)], 2)
        loss = l1(warped_frames, targets)
        (gflows,) = torch.autograd.grad(loss, [flows])
        np.savez_compressed(os.path.join(OUT, name + ".npz"), source=source.numpy(), flows=flows.detach().numpy(),
                            targets=targets.numpy(), loss=loss.detach().numpy(), gflows=gflows.numpy())
        print(name, float(loss.detach()), float(gflows.abs().max()))


if __name__ == "__main__":
    main()
