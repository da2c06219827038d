"""Affine-grid object warp + its objects x T loop (dense_motion.py:94-168) and the flow-consistency loss
(losses.py:115-141).

CPU (-m "not gpu"): the oracle restatements (oracle.reference_torch.affine_warp / generate_sparse_motion /
flow_consistency_loss) against tests/golden/motion/*.npz, which oracle/make_golden_motion.py wrote by running the
UNMODIFIED reference classes.  GPU (-m gpu): the CUDA kernels, through the C ABI, against the oracle run on the same
device (the reference's own CUDA arithmetic -- bit-equal selections are asserted there) and against the fixtures.
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import reference_torch as rt

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "motion", "*.npz")))
WARP = [g for g in GOLD if os.path.basename(g).startswith("warp_")]
SPARSE = [g for g in GOLD if os.path.basename(g).startswith("sparse_")]
FLOWCON = [g for g in GOLD if os.path.basename(g).startswith("flowcon_")]
ids = lambda ps: [os.path.basename(p)[:-4] for p in ps]  # noqa: E731


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    den = b.abs().max().item()
    num = (a - b).abs().max().item()
    return num / den if den > 0 else num


def test_fixtures_exist():
    assert len(WARP) >= 3 and len(SPARSE) >= 3 and len(FLOWCON) >= 2


# ---------------------------------------------------------------------------------------------- oracle vs reference
@pytest.mark.parametrize("path", WARP, ids=ids(WARP))
def test_oracle_affine_warp_matches_reference(path):
    d = np.load(path)
    theta, x = torch.from_numpy(d["theta"]), torch.from_numpy(d["x"])
    base = rt.object_base_grid(x.shape[2], x.shape[3], "cpu")
    for k in range(theta.shape[0]):
        t_x, flow = rt.affine_warp(theta[k], x[k:k + 1], base)
        assert torch.equal(t_x[0], torch.from_numpy(d["t_x"][k]))
        assert torch.equal(flow[0], torch.from_numpy(d["flow"][k]))


@pytest.mark.parametrize("path", SPARSE, ids=ids(SPARSE))
def test_oracle_sparse_motion_matches_reference(path):
    d = np.load(path)
    bw, fw, bn = rt.generate_sparse_motion(torch.from_numpy(d["instance"]), torch.from_numpy(d["ids"]),
                                           torch.from_numpy(d["batch"]), torch.from_numpy(d["thetas"]), int(d["T"]))
    assert torch.equal(bw, torch.from_numpy(d["sparse_motion_bw"]))
    assert torch.equal(fw, torch.from_numpy(d["sparse_motion_fw"]))
    assert torch.equal(bn, torch.from_numpy(d["sparse_motion_bin"]))
    assert bn.sum() > 0 and (bw != 0).any()  # the fixtures do exercise the `== 1` selections


@pytest.mark.parametrize("path", FLOWCON, ids=ids(FLOWCON))
def test_oracle_flow_consistency_matches_reference(path):
    d = np.load(path)
    T = int(d["T"])
    flow = torch.from_numpy(d["flow"]).requires_grad_(True)
    back = torch.from_numpy(d["flowback"]).requires_grad_(True)
    mfw = torch.from_numpy(d["mask_fw"]).requires_grad_(True)
    mbw = torch.from_numpy(d["mask_bw"]).requires_grad_(True)
    v = rt.flow_consistency_loss(flow, back, num_predicted_frames=T)
    g = torch.autograd.grad(v, [flow, back])
    assert torch.equal(v.detach(), torch.from_numpy(d["loss"]))
    assert torch.equal(g[0], torch.from_numpy(d["gflow"])) and torch.equal(g[1], torch.from_numpy(d["gflowback"]))
    vm = rt.flow_consistency_loss(flow, back, mfw, mbw, num_predicted_frames=T)
    gm = torch.autograd.grad(vm, [flow, back, mfw, mbw])
    assert torch.equal(vm.detach(), torch.from_numpy(d["loss_m"]))
    for a, k in zip(gm, ("gflow_m", "gflowback_m", "gmask_fw", "gmask_bw")):
        assert torch.equal(a, torch.from_numpy(d[k]))


# ---------------------------------------------------------------------------------------------------- CUDA kernels
@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda", 0)


@pytest.mark.gpu
@pytest.mark.parametrize("path", WARP, ids=ids(WARP))
def test_affine_warp_kernel(dev, path):
    import c2m_b200
    from c2m_b200 import _lib
    d = np.load(path)
    theta, x = torch.from_numpy(d["theta"]).to(dev), torch.from_numpy(d["x"]).to(dev)
    n0 = _lib.launch_count()
    t_x, flow = c2m_b200.affine_warp(theta, x)
    assert _lib.launch_count() - n0 == 1  # all K matrices in one launch
    # the reference's CUDA arithmetic on the same device: bit for bit
    base = rt.object_base_grid(x.shape[2], x.shape[3], dev)
    for k in range(theta.shape[0]):
        r_tx, r_flow = rt.affine_warp(theta[k], x[k:k + 1], base)
        assert torch.equal(flow[k], r_flow[0]), f"flow {k}: {rel(flow[k], r_flow[0]):.3e}"
        assert torch.equal(t_x[k], r_tx[0]), f"t_x {k}: {rel(t_x[k], r_tx[0]):.3e}"
    # and the CPU-generated fixtures within tolerance (ATen's CPU and CUDA builds round the grid differently)
    assert rel(t_x, d["t_x"]) <= 1e-5 and rel(flow, d["flow"]) <= 1e-5
    # the reference's own call shape: one [2,3] matrix, one image
    a, b = c2m_b200.affine_warp(theta[0], x[0:1], base)
    assert a.shape == (1, x.shape[1], x.shape[2], x.shape[3]) and b.shape == (1, 2, x.shape[2], x.shape[3])
    assert torch.equal(a[0], t_x[0]) and torch.equal(b[0], flow[0])


@pytest.mark.gpu
@pytest.mark.parametrize("path", SPARSE, ids=ids(SPARSE))
def test_sparse_motion_kernel(dev, path):
    import c2m_b200
    from c2m_b200 import _lib
    d = np.load(path)
    T = int(d["T"])
    inst = torch.from_numpy(d["instance"]).to(dev)
    ids_, batch, thetas = (torch.from_numpy(d[k]).to(dev) for k in ("ids", "batch", "thetas"))
    n0 = _lib.launch_count()
    bw, fw, bn = c2m_b200.sparse_motion(inst, ids_, batch, thetas)
    assert _lib.launch_count() - n0 == 1  # the whole objects x T loop
    r_bw, r_fw, r_bn = rt.generate_sparse_motion(inst, ids_, batch, thetas, T)
    # bit-equal to the reference loop run on this device -- including which pixels pass `warped_obj == 1`
    assert torch.equal(bn, r_bn)
    assert torch.equal(bw, r_bw)
    assert torch.equal(fw, r_fw)
    assert bn.sum() > 0
    # fixtures from the CPU run: the selections may differ on a handful of boundary pixels (different grid rounding)
    gb = torch.from_numpy(d["sparse_motion_bin"]).to(dev)
    assert (bn != gb).float().mean().item() < 2e-3
    same = (bn == gb).expand_as(bw)
    assert rel(bw[same], torch.from_numpy(d["sparse_motion_bw"]).to(dev)[same]) <= 1e-5
    assert rel(fw, d["sparse_motion_fw"]) <= 1e-5
    # the drop-in method: same dictionary as the reference's, occlusion maps included
    import types
    me = types.SimpleNamespace(train_params={"num_predicted_frames": T, "use_fw_of": True})
    gnn = types.SimpleNamespace(source_frames_nodes_instance_ids=ids_.view(-1, 1), batch=batch, targets_theta=thetas)
    smd = {f"theta_{t}": thetas[:, t] for t in range(T)}
    out = c2m_b200.generate_sparse_motion(me, gnn, smd, inst)
    out_gt = c2m_b200.generate_sparse_motion(me, gnn, None, inst, use_gt=True)
    assert set(out) == {"sparse_motion_bw", "sparse_motion_fw", "sparse_motion_bin", "sparse_occ_bw", "sparse_occ_fw"}
    assert all(torch.equal(out[k], out_gt[k]) for k in out)
    assert torch.equal(out["sparse_motion_bw"], bw) and torch.equal(out["sparse_motion_bin"], bn)
    for t in range(T):  # dense_motion.py:152-158
        occ_bw = c2m_b200.motion.clip_mask(c2m_b200.get_occlusion_map(fw[:, :, t].contiguous()))
        occ_fw = c2m_b200.motion.clip_mask(c2m_b200.get_occlusion_map(bw[:, :, t].contiguous()))
        assert torch.equal(out["sparse_occ_bw"][:, :, t], occ_bw) and torch.equal(out["sparse_occ_fw"][:, :, t], occ_fw)
    assert rel(out["sparse_occ_bw"], d["sparse_occ_bw"]) <= 0.02 or \
        (out["sparse_occ_bw"].cpu() != torch.from_numpy(d["sparse_occ_bw"])).float().mean() < 2e-3


@pytest.mark.gpu
def test_sparse_motion_larger_scene(dev):
    """128 x 256 (the shipped YAML's size), 3 images x 12 objects x 5 frames: one launch, bit-equal to the loop."""
    import c2m_b200
    from oracle.make_golden_motion import scene
    g = torch.Generator().manual_seed(5)
    inst, ids_, batch, thetas = scene(g, 3, 128, 256, 12, 5)
    inst, ids_, batch, thetas = inst.to(dev), ids_.to(dev), batch.to(dev), thetas.to(dev)
    bw, fw, bn = c2m_b200.sparse_motion(inst, ids_, batch, thetas)
    r_bw, r_fw, r_bn = rt.generate_sparse_motion(inst, ids_, batch, thetas, 5)
    assert torch.equal(bn, r_bn) and torch.equal(bw, r_bw) and torch.equal(fw, r_fw)
    # no objects at all / empty tensors
    e = c2m_b200.sparse_motion(inst, ids_[:0], batch[:0], thetas[:0])
    assert all((t == 0).all() for t in e)


@pytest.mark.gpu
@pytest.mark.parametrize("path", FLOWCON, ids=ids(FLOWCON))
def test_flow_consistency_kernel_vs_fixtures(dev, path):
    import c2m_b200
    d = np.load(path)
    T = int(d["T"])
    mk = lambda k: torch.from_numpy(d[k]).to(dev).requires_grad_(True)  # noqa: E731
    flow, back, mfw, mbw = mk("flow"), mk("flowback"), mk("mask_fw"), mk("mask_bw")
    v = c2m_b200.flow_consistency_loss(flow, back, num_predicted_frames=T)
    g = torch.autograd.grad(v, [flow, back])
    assert rel(v, d["loss"]) <= 1e-5
    assert rel(g[0], d["gflow"]) <= 1e-4 and rel(g[1], d["gflowback"]) <= 1e-4
    vm = c2m_b200.FlowConsistLoss({"num_predicted_frames": T})(flow, back, mfw, mbw)
    gm = torch.autograd.grad(vm, [flow, back, mfw, mbw])
    assert rel(vm, d["loss_m"]) <= 1e-5
    for a, k in zip(gm, ("gflow_m", "gflowback_m", "gmask_fw", "gmask_bw")):
        assert rel(a, d[k]) <= 1e-4, k


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 5, 64, 128), (1, 3, 37, 53), (3, 2, 128, 256)], ids=str)
def test_flow_consistency_kernel_vs_device_oracle(dev, shape):
    import c2m_b200
    B, T, H, W = shape
    g = torch.Generator().manual_seed(sum(shape))
    mk = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).to(dev)  # noqa: E731
    flow, back = mk(B, 2, T, H, W, scale=4.0), mk(B, 2, T, H, W, scale=4.0)
    mfw, mbw = torch.rand(B, 1, T, H, W, generator=g).to(dev), torch.rand(B, 1, T, H, W, generator=g).to(dev)
    for masked in (False, True):
        a = [t.clone().requires_grad_(True) for t in (flow, back, mfw, mbw)]
        b = [t.clone().requires_grad_(True) for t in (flow, back, mfw, mbw)]
        n = 4 if masked else 2
        v = c2m_b200.flow_consistency_loss(a[0], a[1], *(a[2:] if masked else ()), num_predicted_frames=T)
        r = rt.flow_consistency_loss(b[0], b[1], *(b[2:] if masked else ()), num_predicted_frames=T)
        ga = torch.autograd.grad(v * 3.0, a[:n])
        gb = torch.autograd.grad(r * 3.0, b[:n])
        assert rel(v, r) <= 1e-5
        for x, y in zip(ga, gb):
            assert rel(x, y) <= 1e-4
        # bitwise reproducible (the reference's atomicAdd scatter is not)
        a2 = [t.clone().requires_grad_(True) for t in (flow, back, mfw, mbw)]
        v2 = c2m_b200.flow_consistency_loss(a2[0], a2[1], *(a2[2:] if masked else ()), num_predicted_frames=T)
        ga2 = torch.autograd.grad(v2 * 3.0, a2[:n])
        assert torch.equal(v, v2) and all(torch.equal(x, y) for x, y in zip(ga, ga2))
    # only one side needs a gradient
    a = flow.clone().requires_grad_(True)
    v = c2m_b200.flow_consistency_loss(a, back, num_predicted_frames=T)
    b = flow.clone().requires_grad_(True)
    r = rt.flow_consistency_loss(b, back, num_predicted_frames=T)
    assert rel(torch.autograd.grad(v, a)[0], torch.autograd.grad(r, b)[0]) <= 1e-4
