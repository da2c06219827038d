#!/usr/bin/env python
"""BASELINE.json configs[3]: one training step of the occlusion-aware generator (random init, Cityscapes
256x512, 8 clips x 5 frames folded into the batch per GPU) with the fused warp kernels swapped in, next to the
same module running the reference's torch composition (CPU-built grid + H2D + div/cat/add + grid_sample + mul).

    python tools/bench_generator.py [--warp fused|torch|both] [--steps K] [--warmup W] [--frames N]
    python -m torch.distributed.run --nproc-per-node G ... tools/bench_generator.py      (real DDP, NCCL)

Step = forward, MSE against a noise target, backward, Adam.  The convolutions are ordinary cuDNN either way;
the numbers show what the op is worth inside the step, not the op itself (bench.py measures that)."""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from c2m_b200 import dist as cdist  # noqa: E402
from c2m_b200 import generator as cgen  # noqa: E402
from oracle import reference_torch as rt  # noqa: E402  (checker/baseline leg only)

PARAMS = dict(block_expansion=32, num_down_blocks=3, max_expansion=512, num_bottleneck_blocks=4,
              padding_mode="reflect", use_skip=False, use_spade=False)


def run(mode, args, rank, local_rank, world, dev):
    torch.manual_seed(1234)  # same weights on every rank (src/train.py:70)
    net = cgen.OcclusionAwareGenerator(dict(PARAMS), None, 3, args.dataset).to(dev)
    if args.channels_last:
        net = net.to(memory_format=torch.channels_last)
    model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local_rank]) if world > 1 else net
    opt = torch.optim.Adam(net.parameters(), lr=2e-4, betas=(0.5, 0.999))
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    N, H, W = args.frames, 256, 512
    frame = torch.rand(N, 3, H, W, device=dev, generator=g)
    if args.channels_last:
        frame = frame.contiguous(memory_format=torch.channels_last)
    flow = torch.randn(N, 2, H, W, device=dev, generator=g) * 4
    occ = torch.rand(N, 1, H, W, device=dev, generator=g)
    target = torch.rand(N, 3, H, W, device=dev, generator=g)
    cls = cgen.OcclusionAwareGenerator
    saved = (cls.apply_optical, cls.deform_input)
    if mode == "torch":  # the reference composition on the GPU, as the unpatched trainer would run it
        cls.apply_optical = lambda self, input_ref=None, optical_flow=None, occlusion_map=None: rt.apply_optical(
            input_ref, optical_flow, occlusion_map)
        cls.deform_input = staticmethod(rt.deform_input)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = F.mse_loss(model(frame, flow, occ), target)
        loss.backward()
        opt.step()
        return loss

    try:
        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize(dev)
        if world > 1:
            torch.distributed.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / args.steps
    finally:
        cls.apply_optical, cls.deform_input = saved
    value, ms_max, _ = cdist.aggregate_throughput(N, ms, dev)
    return {"frames_per_s": value, "ms_per_step": ms_max, "loss": float(loss.detach())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--warp", default="both", choices=["fused", "torch", "both"])
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=40)
    ap.add_argument("--dataset", default="cityscapes")
    ap.add_argument("--no-channels-last", dest="channels_last", action="store_false")
    args = ap.parse_args()
    saved_stdout = os.dup(1)  # library chatter on stdout (NCCL banner) goes to stderr; the result line to stdout
    os.dup2(2, 1)
    rank, local_rank, world = cdist.init()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    torch.backends.cudnn.benchmark = True  # as the reference trainer sets it (src/train.py:52)
    res = {m: run(m, args, rank, local_rank, world, dev) for m in (["fused", "torch"] if args.warp == "both" else [args.warp])}
    if rank == 0:
        line = {"metric": "generator train step, frames/s (Cityscapes 256x512, fwd+bwd+Adam)", "n_gpus": world,
                "frames_per_gpu": args.frames, "dataset": args.dataset, "channels_last": args.channels_last,
                "parallelism": "ddp%d (NCCL all-reduce of %d parameters)" % (world, 5811459 if args.dataset == "cityscapes" else 7686147),
                "results": res}
        if len(res) == 2:
            line["step_speedup_fused_vs_torch"] = res["torch"]["ms_per_step"] / res["fused"]["ms_per_step"]
        sys.stdout.flush()
        os.write(saved_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
