"""Randomised parity stress (development tool): many random shapes / flows / options, channels-last and NCHW,
against the torch CUDA composition of the reference path.  Prints the worst relative errors; exits 1 on a miss."""
import random
import sys
import os

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import c2m_b200  # noqa: E402

dev = torch.device("cuda", 0)
FORCE_DET = os.environ.get("C2M_STRESS_DET", "0") not in ("", "0")


def ref(x, flow, mask, B):
    N, _, H, W = flow.shape
    xin = x if B == N else x.repeat(N // B, 1, 1, 1)
    g0 = torch.zeros([N, 2, H, W])
    g0[:, 0] = (torch.linspace(-1, 1, W) if W > 1 else torch.Tensor([-1])).view(1, 1, W).expand(N, H, W)
    g0[:, 1] = (torch.linspace(-1, 1, H) if H > 1 else torch.Tensor([-1])).view(1, H, 1).expand(N, H, W)
    g0 = g0.to(dev)
    nf = torch.cat([flow[:, 0:1] / ((W - 1.0) / 2.0), flow[:, 1:2] / ((H - 1.0) / 2.0)], 1)
    o = F.grid_sample(xin, (g0 + nf).permute(0, 2, 3, 1), mode="bilinear", padding_mode="border", align_corners=False)
    return o if mask is None else o * mask


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def main(iters, seed):
    rnd = random.Random(seed)
    worst = {"out": 0.0, "gx": 0.0, "gflow": 0.0, "gmask": 0.0}
    for it in range(iters):
        torch.manual_seed(seed * 100003 + it)
        H, W = rnd.randint(2, 70), rnd.randint(2, 150)
        C = rnd.choice([4, 8, 12, 16, 20, 32, 36, 64, 100, 128, 192, 256, 512, 3, 5])
        T = rnd.choice([1, 1, 1, 2, 5])
        B = rnd.randint(1, 3)
        N = B * T
        kind = rnd.choice(["smooth", "noise", "converge", "oob", "shift"])
        ii = torch.arange(H, device=dev, dtype=torch.float32).view(1, 1, H, 1)
        jj = torch.arange(W, device=dev, dtype=torch.float32).view(1, 1, 1, W)
        flow = torch.randn(N, 2, H, W, device=dev) * rnd.choice([0.3, 1.0, 3.0])
        if kind == "smooth":
            flow = flow * 0.2 + 6 * torch.sin(ii / 9 + jj / 13)
        elif kind == "converge":
            flow = flow + torch.cat([(W / 2 - jj).expand(N, 1, H, W), (H / 2 - ii).expand(N, 1, H, W)], 1) * rnd.choice([0.5, 0.9, 1.0])
        elif kind == "oob":
            flow = flow * W
        elif kind == "shift":
            flow = flow * 0.1 + rnd.choice([-40.0, 17.0, 33.0])
        x = torch.randn(B, C, H, W, device=dev)
        if rnd.random() < 0.7 and C % 4 == 0:
            x = x.contiguous(memory_format=torch.channels_last)
        mask = torch.rand(N, 1, H, W, device=dev) if rnd.random() < 0.8 else None
        gout = torch.randn(N, C, H, W, device=dev)
        det = FORCE_DET or rnd.random() < 0.25
        xs = [x.clone().requires_grad_(True) for _ in range(2)]
        fs = [flow.clone().requires_grad_(True) for _ in range(2)]
        ms = [None if mask is None else mask.clone().requires_grad_(True) for _ in range(2)]
        o1 = c2m_b200.warp_blend(xs[0], fs[0], ms[0], deterministic=det)
        o2 = ref(xs[1], fs[1], ms[1], B)
        ins1 = [t for t in (xs[0], fs[0], ms[0]) if t is not None]
        ins2 = [t for t in (xs[1], fs[1], ms[1]) if t is not None]
        g1 = torch.autograd.grad(o1, ins1, gout, retain_graph=det)
        if det:  # deterministic mode: a second backward of the same graph gives the same bits
            g1b = torch.autograd.grad(o1, ins1, gout)
            if not torch.equal(g1[0], g1b[0]):
                print("NOT REPRODUCIBLE", it, dict(N=N, B=B, C=C, H=H, W=W, kind=kind, cl=not x.is_contiguous()))
                return 1
        g2 = torch.autograd.grad(o2, ins2, gout)
        errs = {"out": rel(o1, o2), "gx": rel(g1[0], g2[0]), "gflow": rel(g1[1], g2[1])}
        if mask is not None:
            errs["gmask"] = rel(g1[2], g2[2])
        for k, v in errs.items():
            worst[k] = max(worst[k], v)
        bad = errs["out"] > 1e-5 or max(errs["gx"], errs["gflow"], errs.get("gmask", 0.0)) > 1e-4
        if bad:
            print("MISS", it, dict(N=N, B=B, C=C, H=H, W=W, kind=kind, det=det, cl=not x.is_contiguous(), mask=mask is not None), errs)
            return 1
    print("ok", iters, "cases; worst", {k: float("%.2e" % v) for k, v in worst.items()})
    # the autograd worker thread may still be releasing the last graph's tensors: see tools/bench_loss_site.py
    torch.cuda.synchronize()
    import time
    time.sleep(0.2)
    return 0


if __name__ == "__main__":
    sys.exit(main(int(sys.argv[1]) if len(sys.argv) > 1 else 300, int(sys.argv[2]) if len(sys.argv) > 2 else 1))
