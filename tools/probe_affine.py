"""GPU probe (development tool): which float32 operation order reproduces torch's CUDA F.affine_grid (align_corners
False) + the reference's flow expression (dense_motion.py:161-168) bit for bit.  Candidates are emulated with float64
intermediates (a product of two float32 is exact in float64, so fma(a,b,c) == float32(float64(a)*b + c) up to a rare
double rounding)."""
import itertools
import sys

import torch
import torch.nn.functional as F

dev = torch.device("cuda", 0)
f32, f64 = torch.float32, torch.float64


def lin_cuda(n):
    return torch.linspace(-1, 1, n, device=dev)


def fma(a, b, c):
    return (a.to(f64) * b.to(f64) + c.to(f64)).to(f32)


def mismatches(a, b):
    return int((a.view(torch.int32) != b.view(torch.int32)).sum().item())


for (H, W) in [(128, 256), (256, 512), (104, 208), (33, 77), (64, 128), (256, 832)]:
    print(f"== {H}x{W}")
    # ---- linspace on CUDA vs the CPU form
    for n in (H, W):
        lc, lcpu = lin_cuda(n), torch.linspace(-1, 1, n).to(dev)
        k = torch.arange(n, device=dev, dtype=f32)
        step = torch.tensor(2.0 / (n - 1), dtype=f32, device=dev)
        one = torch.tensor(1.0, dtype=f32, device=dev)
        c_fma = torch.where(k < n // 2, fma(step, k, -one), fma(-step, (n - 1 - k), one))
        c_nofma = torch.where(k < n // 2, (step * k) - 1.0, 1.0 - step * (n - 1 - k))
        step64 = torch.tensor(2.0 / (n - 1), dtype=f64, device=dev)
        c_d = torch.where(k < n // 2, (-1.0 + step64 * k.to(f64)), (1.0 - step64 * (n - 1 - k).to(f64))).to(f32)
        print(f"  linspace n={n}: cuda vs cpu {mismatches(lc, lcpu)}; vs fma-form {mismatches(lc, c_fma)}; "
              f"vs nofma {mismatches(lc, c_nofma)}; vs double-step {mismatches(lc, c_d)}")
    torch.manual_seed(H * 1000 + W)
    K = 6
    theta = torch.eye(2, 3, device=dev).repeat(K, 1, 1) + 0.15 * torch.randn(K, 2, 3, device=dev)
    grid = F.affine_grid(theta, (K, 1, H, W), align_corners=False)  # [K,H,W,2]
    # ---- base range candidates: linspace * (n-1) / n
    cand_range = {}
    for n in (H, W):
        l = lin_cuda(n)
        cand_range[n] = {
            "mul_then_div": (l * (n - 1)) / n,
            "mul_then_recip": (l * (n - 1)) * torch.tensor(1.0 / n, dtype=f32, device=dev),
            "torch_expr": l * (n - 1) / n,
            "ratio_f32": l * torch.tensor((n - 1) / n, dtype=f32, device=dev),
            "double": (l.to(f64) * (n - 1) / n).to(f32),
        }
    for rx_name, ry_name in itertools.product(cand_range[W], cand_range[H]):
        if rx_name != ry_name:
            continue
        bx = cand_range[W][rx_name].view(1, 1, W).expand(K, H, W)
        by = cand_range[H][ry_name].view(1, H, 1).expand(K, H, W)
        one = torch.ones_like(bx)
        for out_c in range(2):
            t0 = theta[:, out_c, 0].view(K, 1, 1)
            t1 = theta[:, out_c, 1].view(K, 1, 1)
            t2 = theta[:, out_c, 2].view(K, 1, 1).expand(K, H, W)
            cands = {
                "fma012": fma(one, t2, fma(by, t1, (bx * t0))),
                "fma210": fma(bx, t0, fma(by, t1, t2.contiguous())),
                "fma_0_12": fma(bx, t0, fma(by, t1, t2.contiguous() * 0)) + t2,
                "nofma012": ((bx * t0) + (by * t1)) + t2,
                "double": (bx.to(f64) * t0.to(f64) + by.to(f64) * t1.to(f64) + t2.to(f64)).to(f32),
            }
            ref = grid[..., out_c]
            print(f"  range={rx_name:15s} out={out_c}: " + "  ".join(f"{k}={mismatches(v.contiguous(), ref.contiguous())}" for k, v in cands.items()),
                  f"of {ref.numel()}")
    # ---- does our zeros-padding sampler agree with ATen on a binary mask at torch's own grid?
    try:
        sys.path.insert(0, ".")
        import c2m_b200
        from c2m_b200 import _lib
        m = (torch.rand(K, 1, H, W, device=dev) > 0.5).float()
        yy, xx = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
        blob = (((yy - H / 2) ** 2 / (H / 4) ** 2 + (xx - W / 2) ** 2 / (W / 4) ** 2) < 1).float().expand(K, 1, H, W).contiguous()
        for name, mm in (("random", m), ("blob", blob)):
            ref = F.grid_sample(mm, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
            out = torch.empty_like(ref)
            _lib.warp_blend_fwd(mm.data_ptr(), grid.contiguous().data_ptr(), None, None, out.data_ptr(), K, 1, H, W, K,
                                mm.stride(), out.stride(), _lib.PAD_ZEROS, _lib.FLAG_COORD_GRID,
                                torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            print(f"  sampler on torch grid ({name}): bit mismatches {mismatches(out, ref)} of {ref.numel()}; "
                  f"(==1) set differs at {int(((out == 1) != (ref == 1)).sum())}; ones: {int((ref == 1).sum())}")
    except Exception as e:  # noqa: BLE001
        print("  sampler check failed:", repr(e))
