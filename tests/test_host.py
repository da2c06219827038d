"""CPU tests of the host logic: input validation (no CPU fallback), drop-in patching of an imported
reference tree, deterministic-mode selection, batch x frame sharding over a world_size-2 gloo group."""
import os
import sys
import types

import pytest
import torch
import torch.multiprocessing as mp

import c2m_b200
from c2m_b200 import dist as cdist
from c2m_b200 import functional


def test_cpu_tensors_are_rejected_like_the_reference():
    # the reference raises on CPU tensors too (ops.py:189,202: base_grid.cuda(-1))
    x = torch.randn(1, 2, 4, 4)
    f = torch.zeros(1, 2, 4, 4)
    with pytest.raises(RuntimeError):
        c2m_b200.resample(x, f)
    with pytest.raises(RuntimeError):
        c2m_b200.apply_optical(None, x, f, torch.ones(1, 1, 4, 4))


def test_shape_and_dtype_validation():
    chk = functional._check_inputs
    x = torch.empty(2, 3, 4, 5, device="meta")
    with pytest.raises(RuntimeError):
        chk(torch.empty(2, 3, 4, 5), torch.empty(2, 2, 4, 5), None, None)
    # meta tensors are not CUDA either
    with pytest.raises(RuntimeError):
        chk(x, x, None, None)


def test_only_bilinear_mode():
    with pytest.raises(NotImplementedError):
        c2m_b200.resample(torch.zeros(1, 1, 2, 2), torch.zeros(1, 2, 2, 2), mode="nearest")


def test_deterministic_selection(monkeypatch):
    monkeypatch.delenv("C2M_WARP_DETERMINISTIC", raising=False)
    assert functional.deterministic_default() is False
    monkeypatch.setenv("C2M_WARP_DETERMINISTIC", "1")
    assert functional.deterministic_default() is True
    monkeypatch.setenv("C2M_WARP_DETERMINISTIC", "0")
    torch.use_deterministic_algorithms(True)
    try:
        assert functional.deterministic_default() is True
    finally:
        torch.use_deterministic_algorithms(False)


def test_nchw_policy_and_plan_selection(monkeypatch):
    """Host policies of c2m_b200/functional.py, decided from tensor metadata and the environment only."""
    from c2m_b200 import _lib
    monkeypatch.delenv("C2M_WARP_NCHW", raising=False)
    feat = torch.empty(2, 64, 8, 16, device="meta")
    assert functional._promotes(feat, 0)                                    # feature maps are converted once
    assert not functional._promotes(torch.empty(2, 3, 8, 16, device="meta"), 0)    # image-like tensors are not
    assert not functional._promotes(torch.empty(2, 10, 8, 16, device="meta"), 0)   # nor C % 4 != 0
    assert not functional._promotes(torch.empty(0, 64, 8, 16, device="meta"), 0)
    for flag in (_lib.FLAG_STRICT_LAYOUT, _lib.FLAG_NO_STAGE, _lib.FLAG_FORCE_GENERIC, _lib.FLAG_COORD_GRID,
                 _lib.FLAG_BWD_ATOMIC, _lib.FLAG_ALIGN_CORNERS):
        assert not functional._promotes(feat, flag)
    monkeypatch.setenv("C2M_WARP_NCHW", "strict")
    assert not functional._promotes(feat, 0) and functional._relayout_ok(feat)
    monkeypatch.setattr(functional, "_STAGE_MAX_BYTES", 1024)
    assert not functional._relayout_ok(feat)                                # the copy would exceed the budget
    # the plan is opt-in
    monkeypatch.delenv("C2M_WARP_PLAN", raising=False)
    assert functional._plan_enabled() is False
    monkeypatch.setenv("C2M_WARP_PLAN", "1")
    assert functional._plan_enabled() is True
    # c2m_warp_plan_bytes is a pure function of its arguments (no GPU): 24 B per output pixel and a head, or 0
    n = _lib.plan_bytes(40, 64, 256, 512, 40, 0)
    assert 22 * 40 * 256 * 512 < n < 30 * 40 * 256 * 512
    assert _lib.plan_bytes(40, 64, 256, 512, 40, _lib.FLAG_DETERMINISTIC) == 0
    assert _lib.plan_bytes(40, 64, 256, 512, 40, _lib.FLAG_FORCE_GENERIC) == 0
    assert _lib.plan_bytes(40, 66, 256, 512, 40, 0) == 0                    # C % 4 != 0: no channels-last gather
    assert _lib.plan_bytes(40, 64, 256, 512, 7, 0) == 0                     # x_batch must divide N
    assert _lib.plan_bytes(70000, 64, 8, 32, 70000, 0) == 0                 # beyond the gather's frame limit
    # argument errors of the relayout entry point are reported before anything is launched
    with pytest.raises(_lib.C2MWarpError):
        _lib.relayout(0, 0, 70000, 8, 4, 4, True, None)
    with pytest.raises(_lib.C2MWarpError):
        _lib.relayout(None, None, 2, 8, 4, 4, True, None)
    _lib.relayout(None, None, 0, 8, 4, 4, True, None)                       # empty: nothing to do


def test_patch_reference_rebinds_every_bound_name(monkeypatch):
    fake = {}
    for name in ("utils", "utils.ops", "modules", "modules.generator", "modules.generator.generator",
                 "modules.motion_estimator", "modules.motion_estimator.motion_autoencoder", "losses",
                 "losses.losses"):
        fake[name] = types.ModuleType(name)
        monkeypatch.setitem(sys.modules, name, fake[name])
    sentinel = object()
    for m in ("utils", "utils.ops"):
        fake[m].resample = fake[m].grid_sample = fake[m].get_grid = sentinel
        fake[m].get_occlusion_map = fake[m].get_corresponding_map = sentinel
    for m in ("modules.generator.generator", "modules.motion_estimator.motion_autoencoder", "losses.losses"):
        fake[m].resample = sentinel

    class OcclusionAwareGenerator:  # stands in for generator.py:11
        @staticmethod
        def deform_input(inp, flow):
            return sentinel

        def apply_optical(self, input_ref=None, optical_flow=None, occlusion_map=None):
            return sentinel

    fake["modules.generator.generator"].OcclusionAwareGenerator = OcclusionAwareGenerator
    done = c2m_b200.patch_reference()
    assert len(done) == 14
    assert fake["utils"].get_occlusion_map is c2m_b200.get_occlusion_map
    assert fake["utils.ops"].resample is c2m_b200.resample
    assert fake["utils"].get_grid is c2m_b200.get_grid
    assert fake["losses.losses"].resample is c2m_b200.resample
    assert OcclusionAwareGenerator.deform_input is c2m_b200.deform_input
    g = OcclusionAwareGenerator()
    with pytest.raises(RuntimeError):  # reaches our op, which refuses CPU tensors
        g.apply_optical(input_ref=torch.zeros(1, 1, 2, 2), optical_flow=torch.zeros(1, 2, 2, 2))


def test_shard_range_partitions_exactly():
    for n in (0, 1, 5, 40, 64, 67):
        for world in (1, 2, 3, 4, 8):
            spans = [cdist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        cdist.shard_range(4, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, lr, w = cdist.init(backend="gloo")
    begin, end = cdist.shard_range(41, r, w)
    # rank 1 is "slower": whole-job throughput must use the max time and the summed frames
    fps, ms, total = cdist.aggregate_throughput(end - begin, 10.0 * (1 + r))
    q.put((r, begin, end, fps, ms, total))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_two_rank_gloo_sharding_and_timing_aggregation():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, b0, e0, fps0, ms0, t0), (r1, b1, e1, fps1, ms1, t1) = res
    assert (b0, e0, b1, e1) == (0, 21, 21, 41)
    assert t0 == t1 == 41 and ms0 == ms1 == 20.0
    assert abs(fps0 - 41 / 0.020) < 1e-6 and fps0 == fps1
