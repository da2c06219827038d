"""The host module c2m_b200.generator.OcclusionAwareGenerator against the unmodified reference class
(/root/reference/src/modules/generator/generator.py) -- CPU, only where the reference tree is present -- and,
on the GPU, with the fused kernels against the same module running the reference's torch composition."""
import os
import sys
import types

import pytest
import torch

import c2m_b200
from c2m_b200 import generator as cgen
from oracle import reference_torch as rt

REF_SRC = "/root/reference/src"
PARAMS = dict(block_expansion=32, num_down_blocks=3, max_expansion=512, num_bottleneck_blocks=4,
              padding_mode="reflect", use_skip=False, use_spade=False)


def _oracle_warp(monkeypatch):
    """Run our module with the oracle's torch composition instead of the CUDA kernels."""
    # (apply_optical hands the flow / mask over at their own sizes with flow_resize="half_pixel": the oracle's
    # apply_optical is the reference composition of that, generator.py:80-96)
    monkeypatch.setattr(cgen, "warp_blend", lambda x, f, m=None, *a, flow_resize=None, **k:
                        rt.apply_optical(x, f, m) if flow_resize == "half_pixel" else rt.warp_blend(x, f, m))
    monkeypatch.setattr(cgen, "resample", rt.resample)


@pytest.fixture()
def ref_class(monkeypatch):
    if not os.path.isdir(REF_SRC):
        pytest.skip("reference tree not present (GPU box)")
    monkeypatch.setitem(sys.modules, "imageio", types.ModuleType("imageio"))
    monkeypatch.syspath_prepend(REF_SRC)
    import modules.generator.generator as refgen
    monkeypatch.setattr(refgen, "resample", rt.resample)  # the reference's own resample needs a GPU (ops.py:189)
    return refgen.OcclusionAwareGenerator


@pytest.mark.parametrize("dataset", ["cityscapes", "kitti"])
def test_state_dict_and_forward_match_the_reference(ref_class, monkeypatch, dataset):
    torch.manual_seed(0)
    ref = ref_class(dict(PARAMS), None, 3, dataset)
    ours = cgen.OcclusionAwareGenerator(dict(PARAMS), None, 3, dataset)
    sd = ref.state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    ours.load_state_dict(sd, strict=True)  # a reference checkpoint loads unchanged
    _oracle_warp(monkeypatch)
    frame = torch.rand(2, 3, 64, 128)
    flow = torch.randn(2, 2, 64, 128) * 3
    occ = torch.rand(2, 1, 64, 128)
    for train in (True, False):
        ref.train(train)
        ours.train(train)
        a = ref(frame, flow, occ)
        b = ours(frame, flow, occ)
        assert a.shape == (2, 3, 64, 128)
        assert torch.allclose(a, b, rtol=0, atol=1e-6), (a - b).abs().max()
    # gradients through the whole module
    ga = torch.autograd.grad(ref(frame, flow, occ).square().mean(), list(ref.parameters()))
    gb = torch.autograd.grad(ours(frame, flow, occ).square().mean(), list(ours.parameters()))
    for x, y in zip(ga, gb):
        assert torch.allclose(x, y, rtol=1e-4, atol=1e-7)


def test_spade_variant_is_refused():
    with pytest.raises(NotImplementedError):
        cgen.OcclusionAwareGenerator(dict(PARAMS, use_spade=True), None, 3, "cityscapes")


@pytest.mark.gpu
@pytest.mark.parametrize("dataset", ["cityscapes", "kitti"])
@pytest.mark.parametrize("channels_last", [False, True])
def test_module_on_gpu_fused_vs_torch_composition(monkeypatch, dataset, channels_last):
    dev = torch.device("cuda", 0)
    torch.manual_seed(1)
    # TF32 convolutions round their inputs to 10 mantissa bits: a 1e-7 difference in the warp output can flip
    # such a rounding and show up as 1e-4 downstream -- compare in full fp32
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    net = cgen.OcclusionAwareGenerator(dict(PARAMS), None, 3, dataset).to(dev).eval()
    if channels_last:
        net = net.to(memory_format=torch.channels_last)
    frame = torch.rand(4, 3, 128, 256, device=dev)
    if channels_last:
        frame = frame.contiguous(memory_format=torch.channels_last)
    flow = (torch.randn(4, 2, 128, 256, device=dev) * 4).requires_grad_(True)
    occ = torch.rand(4, 1, 128, 256, device=dev).requires_grad_(True)
    out = net(frame, flow, occ)
    g = torch.autograd.grad(out.square().mean(), [flow, occ] + list(net.parameters()))
    with monkeypatch.context() as mp:
        mp.setattr(cgen, "warp_blend", lambda x, f, m=None, *a, flow_resize=None, **k:
                   rt.apply_optical(x, f, m) if flow_resize == "half_pixel" else rt.warp_blend(x, f, m))
        mp.setattr(cgen, "resample", rt.resample)
        ref = net(frame, flow, occ)
        gr = torch.autograd.grad(ref.square().mean(), [flow, occ] + list(net.parameters()))
    rel = lambda a, b: ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()  # noqa: E731
    assert rel(out, ref) <= 1e-5
    # (a bias in front of a norm layer has a mathematically zero gradient: pure rounding noise on both sides,
    # hence the small absolute term)
    floor = 1e-6 * max(b.abs().max().item() for b in gr)
    bad = [(k, (a - b).abs().max().item(), b.abs().max().item()) for k, (a, b) in enumerate(zip(g, gr))
           if (a - b).abs().max() > 1e-4 * b.abs().max() + floor]
    assert not bad, bad
