"""TEST INFRASTRUCTURE -- generates tests/golden/motion/*.npz by running the UNMODIFIED reference code on the CPU of the
build container:

  * ``DenseMotionNetwork.warp`` and ``DenseMotionNetwork.generate_sparse_motion``
    (/root/reference/src/modules/motion_estimator/dense_motion.py:94-168), called unbound with a stand-in ``self`` that
    carries ``train_params`` and the class's own static methods;
  * ``FlowConsistLoss`` (/root/reference/src/losses/losses.py:115-141) with the module's ``resample`` name rebound to the
    oracle's device-fixed restatement (the reference's own needs a GPU, ops.py:189,202).

``imageio`` and ``torch_geometric`` are absent from this image; neither is used by the functions exercised here, so
they are stubbed at import time (the sibling modules import them by name).

Run:  python oracle/make_golden_motion.py      (needs /root/reference; the GPU box never runs this)
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/src"
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
OUT = os.path.join(ROOT, "tests", "golden", "motion")
sys.path.insert(0, ROOT)


class _Stub:
    def __init__(self, *a, **k):
        pass


class _AnyModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Stub


def import_reference():
    sys.modules.setdefault("imageio", types.ModuleType("imageio"))
    for n in ("torch_geometric", "torch_geometric.nn", "torch_geometric.utils", "torch_geometric.data",
              "torch_geometric.nn.conv", "torch_geometric.nn.inits", "torch_scatter", "torch_sparse"):
        sys.modules.setdefault(n, _AnyModule(n))
    sys.path.insert(0, REF)
    import modules.motion_estimator.dense_motion as dm
    return dm


def scene(g, B, H, W, n_per_image, T):
    """A synthetic instance map (rectangles / ellipses with ids 26000 + k over background 0, as the dataset's
    instance ids are large integers stored in a float tensor) and near-identity affine motions."""
    inst = torch.zeros(B, 1, H, W)
    ids, batch = [], []
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    for b in range(B):
        for k in range(n_per_image):
            cy, cx = torch.rand(2, generator=g).tolist()
            ry, rx = (0.08 + 0.2 * torch.rand(2, generator=g)).tolist()
            oid = float(26000 + 1000 * b + k)
            if k % 2:
                m = ((yy - cy * H).abs() < ry * H) & ((xx - cx * W).abs() < rx * W)
            else:
                m = ((yy - cy * H) / (ry * H)) ** 2 + ((xx - cx * W) / (rx * W)) ** 2 < 1
            inst[b, 0][m] = oid
            ids.append(oid)
            batch.append(b)
        ids.append(0.0)  # a node with instance id 0 is skipped (dense_motion.py:126-127)
        batch.append(b)
    n = len(ids)
    thetas = torch.eye(2, 3).view(1, 1, 6).repeat(n, T, 1) + 0.08 * torch.randn(n, T, 6, generator=g)
    thetas[0, 0] = torch.tensor([1.0, 0.0, 0.0, 0.0, 1.0, 0.0])           # identity
    thetas[1 % n, 0] = torch.tensor([1.0, 0.0, 4.0 / W, 0.0, 1.0, -2.0 / H])  # integer-pixel translation
    return inst, torch.tensor(ids), torch.tensor(batch, dtype=torch.long), thetas


def main():
    dm = import_reference()
    from oracle import reference_torch as rt
    os.makedirs(OUT, exist_ok=True)
    g = torch.Generator().manual_seed(20261020)
    cls = dm.DenseMotionNetwork
    # ---- warp(): arbitrary images
    for name, (C, H, W) in {"warp_c3_16x24": (3, 16, 24), "warp_c1_33x77": (1, 33, 77), "warp_c2_64x128": (2, 64, 128)}.items():
        K = 4
        theta = torch.eye(2, 3).repeat(K, 1, 1) + 0.2 * torch.randn(K, 2, 3, generator=g)
        x = torch.randn(K, C, H, W, generator=g)
        base = rt.object_base_grid(H, W, "cpu")
        tx, fl = [], []
        for k in range(K):
            a, b = cls.warp(theta[k], x[k:k + 1], base)
            tx.append(a)
            fl.append(b)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), kind="warp", theta=theta.numpy(), x=x.numpy(),
                            t_x=torch.cat(tx).numpy(), flow=torch.cat(fl).contiguous().numpy())
    # ---- generate_sparse_motion(): the object loop
    for name, (B, H, W, nobj, T) in {"sparse_2x32x64": (2, 32, 64, 3, 2), "sparse_3x64x128": (3, 64, 128, 5, 5),
                                     "sparse_1x37x53": (1, 37, 53, 4, 3)}.items():
        inst, ids, batch, thetas = scene(g, B, H, W, nobj, T)
        me = types.SimpleNamespace(train_params={"num_predicted_frames": T, "use_fw_of": True}, warp=cls.warp,
                                   clip_mask=cls.clip_mask)
        gnn = types.SimpleNamespace(source_frames_nodes_instance_ids=ids.view(-1, 1), batch=batch, targets_theta=thetas)
        smd = {f"theta_{t}": thetas[:, t] for t in range(T)}
        out = cls.generate_sparse_motion(me, gnn, smd, inst, use_gt=False)
        out_gt = cls.generate_sparse_motion(me, gnn, smd, inst, use_gt=True)
        assert all(torch.equal(out[k], out_gt[k]) for k in out)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), kind="sparse", instance=inst.numpy(), ids=ids.numpy(),
                            batch=batch.numpy(), thetas=thetas.numpy(), T=T,
                            **{k: v.numpy() for k, v in out.items()})
    # ---- FlowConsistLoss
    import losses.losses as ref_losses
    ref_losses.resample = rt.resample
    for name, (B, T, H, W) in {"flowcon_2x3x16x24": (2, 3, 16, 24), "flowcon_1x5x32x64": (1, 5, 32, 64)}.items():
        flow = (torch.randn(B, 2, T, H, W, generator=g) * 2).requires_grad_(True)
        back = (torch.randn(B, 2, T, H, W, generator=g) * 2).requires_grad_(True)
        mfw = torch.rand(B, 1, T, H, W, generator=g).requires_grad_(True)
        mbw = torch.rand(B, 1, T, H, W, generator=g).requires_grad_(True)
        mod = ref_losses.FlowConsistLoss({"num_predicted_frames": T})
        rec = {"kind": "flowcon", "flow": flow.detach().numpy(), "flowback": back.detach().numpy(),
               "mask_fw": mfw.detach().numpy(), "mask_bw": mbw.detach().numpy(), "T": T}
        v = mod(flow, back)
        gf, gb = torch.autograd.grad(v, [flow, back])
        rec.update(loss=v.detach().numpy(), gflow=gf.numpy(), gflowback=gb.numpy())
        vm = mod(flow, back, mfw, mbw)
        gs = torch.autograd.grad(vm, [flow, back, mfw, mbw])
        rec.update(loss_m=vm.detach().numpy(), gflow_m=gs[0].numpy(), gflowback_m=gs[1].numpy(),
                   gmask_fw=gs[2].numpy(), gmask_bw=gs[3].numpy())
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
