#!/usr/bin/env python
"""bench.py -- warped frames/s, forward+backward, of the fused flow-warp + occlusion-blend op.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one forward + one backward (grad-input, grad-flow, grad-mask) of the op over one batch
of synthetic Cityscapes-shaped frames (BASELINE.json configs[1]: 8 clips x 5 frames = 40 frames,
C=64, 256x512, fp32; channels-last memory format by default -- the layout the sm_100a kernels are
built around -- with the NCHW-contiguous figure reported beside it).  A frame is one [C,H,W] slice of the folded batch x frame axis.

Prints ONE JSON line (rank 0).  Keys follow the driver's contract; see DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (frames per GPU, C, H, W, oob flow)
    "cityscapes_256x512_c64": (40, 64, 256, 512, False),   # BASELINE.json configs[1] headline level
    "kitti_256x832_c64_oob": (40, 64, 256, 832, True),     # configs[2]
    "fullres_1024x2048_c256": (8, 256, 1024, 2048, False),  # configs[4]
    "cpu_128x256_c64": (5, 64, 128, 256, False),           # configs[0]
}
METRIC = "warped frames/sec fwd+bwd @256x512 (fused flow-warp + occlusion blend, fp32)"


def synth(N, C, H, W, oob, seed, device, pin=False):
    """Synthetic inputs of SURVEY.md section 8d: x ~ N(0,1); flow = 8 px low-frequency field +
    N(0,1) px noise (or the large / out-of-bounds variant); mask = sigmoid(N(0,1)); gout ~ N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    big = N * C * H * W >= 2 ** 31 and not pin and torch.device(device).type == "cuda"
    if big:
        # full-resolution shards (17 GB per tensor): generate the two large tensors on the device instead of
        # holding them in host memory once per rank
        gd = torch.Generator(device=device).manual_seed(seed)
        x = torch.randn(N, C, H, W, generator=gd, device=device)
    else:
        x = torch.randn(N, C, H, W, generator=g)
    ii = torch.arange(H, dtype=torch.float32).view(1, H, 1)
    jj = torch.arange(W, dtype=torch.float32).view(1, 1, W)
    two_pi = 6.283185307179586
    fx = 8.0 * torch.sin(two_pi * ii / (H / 2.0)) * torch.cos(two_pi * jj / (W / 2.0))
    fy = 8.0 * torch.cos(two_pi * ii / (H / 2.0)) * torch.sin(two_pi * jj / (W / 2.0))
    noise = float(os.environ.get("C2M_BENCH_FLOW_NOISE", "1.0"))  # px, SURVEY.md 8d uses 1.0
    flow = torch.stack([fx.expand(N, H, W), fy.expand(N, H, W)], 1) + noise * torch.randn(N, 2, H, W, generator=g)
    if oob:
        flow = torch.randn(N, 2, H, W, generator=g) * (W / 4.0)
        sel = torch.rand(N, 1, H, W, generator=g) < 0.05
        flow = torch.where(sel, torch.sign(flow) * 10.0 * W, flow)
    mask = torch.sigmoid(torch.randn(N, 1, H, W, generator=g))
    gout = torch.randn(N, C, H, W, generator=gd, device=device) if big else torch.randn(N, C, H, W, generator=g)
    ts = [x, flow.contiguous(), mask, gout]
    if pin:
        return [t.pin_memory() for t in ts]
    return [t.to(device) for t in ts]


def fwd_bytes(N, C, H, W):
    return 4 * N * H * W * (2 * C + 3)


def bwd_bytes(N, C, H, W):
    return 4 * N * H * W * (3 * C + 6)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of the GPUs in use, sampled for as long as the benchmark runs (one
    nvidia-smi process; it takes over a second to come up on an 8-GPU box, so it is started once).  Rows are
    time-stamped; `window(t0, t1)` summarises the rows that fall inside a timed region, `begin()` / `end()` /
    `summary()` do the same for the main one."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, indices):
        self.indices = list(indices)
        self.rows, self.proc, self.t, self.t0, self.t1 = [], None, None, None, None

    def __enter__(self):
        if not self.indices:
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", ",".join(str(i) for i in self.indices),
                                          "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            deadline = time.perf_counter() + 15.0
            while not self.rows and time.perf_counter() < deadline and self.proc.poll() is None:
                time.sleep(0.01)
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def begin(self):
        self.t0 = time.perf_counter()

    def end(self):
        self.t1 = time.perf_counter()

    def __exit__(self, *a):
        # the sampler is torn down completely (process reaped, reader thread joined, pipe closed): nothing of it
        # is left for the interpreter's shutdown to trip over
        if self.proc is not None:
            time.sleep(0.1)  # one more sampling period: a short region still gets the row that closes it
            self.proc.terminate()
            try:
                self.proc.wait(5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
                self.proc.wait()
            if self.t is not None:
                self.t.join(5)
            self.proc.stdout.close()
            self.proc = None

    def summary(self):
        return self.window(self.t0, self.t1)

    def window(self, t0, t1):
        if not self.indices:
            return None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0, t1 = t0 or 0.0, (t1 or float("inf")) + 0.05
        rows = list(self.rows)
        inside = [r for (t, r) in rows if t0 <= t <= t1]
        widened = False
        if not inside:  # region shorter than one sampling period of a many-GPU query: nearest rows instead
            inside = [r for (t, r) in rows if t0 - 0.3 <= t <= t1 + 0.3]
            widened = True
        sm, mx, reasons = [], 0.0, set()
        for r in inside:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for k, nme in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        busy = [v for v in sm if v > 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "gpus": len(self.indices), **({"window": "+-0.3 s"} if widened else {})}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def workload_config(workload, N, C, H, W, det):
    """The `config` object of the JSON line.  It describes the WORKLOAD and is the same for both arms (`--impl
    ours` and `--impl reference` run the same frames); what differs between the arms -- memory format, sample
    size of the CPU run -- is reported next to it, not inside it."""
    return {"workload": workload, "frames_per_gpu": N, "C": C, "H": H, "W": W,
            "grads": "input+flow+mask", "deterministic": bool(det),
            "l2": "inputs (%.0f MB per tensor) larger than L2, no flush needed" % (4e-6 * N * C * H * W),
            "partition": "batch x frame, %d frames per rank, no collective" % N}


def cpu_reference_run(workload, frames, steps, warmup, threads=None, forward_only=False):
    """The reference's CPU path (oracle.reference_torch: ops.py:187-202 + generator.py:93 restated
    with the single device fix) on `frames` frames of the workload, all host threads: fwd+bwd, or the forward
    alone (BASELINE.json configs[0])."""
    from oracle import reference_torch as rt
    _, C, H, W, oob = WORKLOADS[workload]
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    x, flow, mask, gout = synth(frames, C, H, W, oob, 1234, "cpu")
    if not forward_only:
        x.requires_grad_(True)
        flow.requires_grad_(True)
        mask.requires_grad_(True)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        if forward_only:
            with torch.no_grad():
                rt.warp_blend(x, flow, mask)
        else:
            out = rt.warp_blend(x, flow, mask)
            torch.autograd.grad(out, [x, flow, mask], gout)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return times, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    workload = args.workload
    N, C, H, W, _ = WORKLOADS[workload]
    frames = args.cpu_frames or args.frames or N  # the whole step of the product arm unless told otherwise
    times, threads = cpu_reference_run(workload, frames, args.steps, max(args.warmup, 1))
    total = sum(times)
    value = frames * len(times) / total
    sample = f"{frames} frames (C={C}, {H}x{W}) fwd+bwd per step, {len(times)} steps, torch {torch.__version__} CPU"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(workload, frames, C, H, W, False),
        "layout": "nchw",
        "note": "reference CPU path (oracle.reference_torch: the reference's own python restated with its one "
                "device fix) on the host cores, same frames per step as the product arm",
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------------------------
def _time_steps(fn, steps, warmup, barrier):
    """`steps` calls of fn between two CUDA events on the current stream (after `warmup` untimed calls), barrier +
    synchronize on both sides; returns (ms per call, host perf_counter at begin, at end)."""
    for _ in range(warmup):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier()
    return e0.elapsed_time(e1) / steps, t0, time.perf_counter()


def measure_config(name, dev, rank, world, clk, barrier, peak, nhwc=True, det=False, steps=5, warmup=3, blend=False):
    """One secondary workload (BASELINE.json configs[2] / configs[4], or the north-star blend operand) measured
    outside the main timed region: fwd+bwd over its own tensors, its own clock window."""
    import c2m_b200
    from c2m_b200 import dist as cdist
    N, C, H, W, oob = WORKLOADS[name]
    x, flow, mask, gout = synth(N, C, H, W, oob, 4321 + rank, dev)
    other = None
    if nhwc:
        x = x.contiguous(memory_format=torch.channels_last)
        gout = gout.contiguous(memory_format=torch.channels_last)
    if blend:
        other = torch.randn_like(x).requires_grad_(True)
    x.requires_grad_(True)
    flow.requires_grad_(True)
    mask.requires_grad_(True)
    ins = [x, flow, mask] + ([other] if blend else [])
    evs = []

    def step():
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        out = c2m_b200.warp_blend(x, flow, mask, other, deterministic=det)
        ev[1].record()
        torch.autograd.grad(out, ins, gout)
        ev[2].record()
        evs.append(ev)

    ms, t0, t1 = _time_steps(step, steps, warmup, barrier)
    evs = evs[-steps:]
    fwd_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in evs)
    bwd_ms = statistics.mean(e[1].elapsed_time(e[2]) for e in evs)
    fb, bb = fwd_bytes(N, C, H, W), bwd_bytes(N, C, H, W)
    if blend:  # + read `other` (fwd), + write grad-other (bwd)
        fb += 4 * N * C * H * W
        bb += 4 * N * C * H * W
    value, ms_max, _ = cdist.aggregate_throughput(N * steps, ms * steps, dev)
    res = {"frames_per_gpu": N, "C": C, "H": H, "W": W, "layout": "nhwc" if nhwc else "nchw", "deterministic": det,
           "flow": "large / out-of-bounds (sigma W/4, 5% at +-10 W)" if oob else "smooth 8 px + N(0,1) px",
           "steps": steps, "ms_per_step": ms_max / steps, "frames_per_s": value,
           "fwd": {"ms": fwd_ms, "achieved": fb / (fwd_ms * 1e-3) / 1e9, "frac": fb / (fwd_ms * 1e-3) / 1e9 / peak},
           "bwd": {"ms": bwd_ms, "achieved": bb / (bwd_ms * 1e-3) / 1e9, "frac": bb / (bwd_ms * 1e-3) / 1e9 / peak},
           "achieved": (fb + bb) / (ms * 1e-3) / 1e9, "frac": (fb + bb) / (ms * 1e-3) / 1e9 / peak,
           "clocks": clk.window(t0, t1)}
    if blend:
        res["operand"] = "other: out = m*warp + (1-m)*other, grads input+flow+mask+other"
    del x, flow, mask, gout, other, ins, evs
    torch.cuda.empty_cache()
    return res


def measure_traffic(args, kernel_regex):
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of ONE launch of the dominant kernel, measured in
    this run: ncu profiles a child process that runs one step of the same workload (outside every timed region).
    Returns (bytes or None, note)."""
    import csv
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, "ncu not found"
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none",
           "-k", "regex:" + kernel_regex, "-s", "2", "-c", "1", "--csv",
           sys.executable, os.path.abspath(__file__), "--traffic-child", "--workload", args.workload,
           "--layout", args.layout, "--flags", args.flags] + (["--deterministic"] if args.deterministic else []) + (
               ["--frames", str(args.frames)] if args.frames else [])
    try:
        env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=150, env=env)
    except (OSError, subprocess.TimeoutExpired) as e:
        return None, "ncu child failed: %s" % type(e).__name__
    total, seen = 0.0, 0
    for row in csv.reader(r.stdout.splitlines()):
        if len(row) > 5 and row[-3] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            try:
                v = float(row[-1].replace(",", ""))
            except ValueError:
                continue
            unit = row[-2].lower()
            v *= {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
            total += v
            seen += 1
    if seen != 2:
        return None, "ncu child gave no metrics (rc %d): %s" % (r.returncode, (r.stderr or r.stdout)[-200:].replace("\n", " "))
    return int(total), "ncu dram__bytes_read.sum + dram__bytes_write.sum of one launch, captured in this run (child process)"


def run_traffic_child(args):
    """Child of measure_traffic(): three steps of the workload, nothing else (ncu picks one launch)."""
    import c2m_b200
    dev = torch.device("cuda", 0)
    N, C, H, W, oob = WORKLOADS[args.workload]
    N = args.frames or N
    x, flow, mask, gout = synth(N, C, H, W, oob, 1234, dev)
    if args.layout == "nhwc":
        x = x.contiguous(memory_format=torch.channels_last)
        gout = gout.contiguous(memory_format=torch.channels_last)
    for t in (x, flow, mask):
        t.requires_grad_(True)
    for _ in range(3):
        out = c2m_b200.warp_blend(x, flow, mask, deterministic=bool(args.deterministic), flags=int(args.flags, 0))
        torch.autograd.grad(out, [x, flow, mask], gout)
    torch.cuda.synchronize()
    return 0


def run_ours(args):
    import c2m_b200
    from c2m_b200 import _lib
    from c2m_b200 import dist as cdist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    rank, local_rank, world = cdist.init()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _lib.load()
    workload = args.workload
    N, C, H, W, oob = WORKLOADS[workload]
    if args.frames:
        N = args.frames
    nhwc = args.layout == "nhwc"
    x, flow, mask, gout = synth(N, C, H, W, oob, 1234 + rank, dev)
    if nhwc:
        x = x.contiguous(memory_format=torch.channels_last)
        gout = gout.contiguous(memory_format=torch.channels_last)
    x.requires_grad_(True)
    flow.requires_grad_(True)
    mask.requires_grad_(True)
    det = bool(args.deterministic)
    flags = int(args.flags, 0)
    peak, peak_src = measured_peak()
    if not nhwc:
        # the upstream gradient arrives in the format of the result it belongs to (an NCHW x gives channels-last
        # strided results unless the strict policy is on: c2m_b200/functional.py)
        with torch.no_grad():
            o_probe = c2m_b200.warp_blend(x.detach(), flow.detach(), mask.detach(), deterministic=det, flags=flags)
        if not o_probe.is_contiguous():
            gout = gout.contiguous(memory_format=torch.channels_last)
        del o_probe

    def step(ev=None):
        if ev:
            ev[0].record()
        out = c2m_b200.warp_blend(x, flow, mask, deterministic=det, flags=flags)
        if ev:
            ev[1].record()
        g = torch.autograd.grad(out, [x, flow, mask], gout)
        if ev:
            ev[2].record()
        return out, g

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _lib.launch_count()
    # rank 0 samples the clocks of every GPU of the job (the ranks use GPUs 0 .. world-1 of the box) for the whole
    # run; every timed region below gets the rows of its own window
    with ClockSampler(range(world) if rank == 0 else []) as clk:
        barrier()
        clk.begin()
        t_begin.record()
        for k in range(args.steps):
            step(evs[k])
        t_end.record()
        barrier()
        clk.end()
        launches = _lib.launch_count() - launches0
        ms_total = t_begin.elapsed_time(t_end)
        fwd_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in evs)
        bwd_ms = statistics.mean(e[1].elapsed_time(e[2]) for e in evs)
        value, ms_max, total_frames = cdist.aggregate_throughput(N * args.steps, ms_total, dev)

        # ---- secondary figure: the same step in the other memory format (not part of the timed region)
        other = None
        if not args.no_other_layout:
            if nhwc:
                x2, g2 = x.detach().contiguous().requires_grad_(True), gout.contiguous()
            else:
                x2 = x.detach().contiguous(memory_format=torch.channels_last).requires_grad_(True)
                g2 = gout.contiguous(memory_format=torch.channels_last)

            def time_other(flags2, g):
                def step2():
                    o = c2m_b200.warp_blend(x2, flow, mask, deterministic=det, flags=flags | flags2)
                    torch.autograd.grad(o, [x2, flow, mask], g)

                oms, ot0, ot1 = _time_steps(step2, 5, 3, barrier)
                oach = (fwd_bytes(N, C, H, W) + bwd_bytes(N, C, H, W)) / (oms * 1e-3) / 1e9
                return {"ms_per_step": oms, "frames_per_s_per_gpu": N / (oms * 1e-3), "achieved": oach,
                        "frac": oach / peak, "clocks": clk.window(ot0, ot1)}

            if nhwc:
                # NCHW-contiguous x, the reference's layout.  Default policy: x is converted once in the forward and
                # the results come back channels-last strided, so the upstream gradient arrives in that format too
                # (what a consumer of a channels_last tensor hands back); "strict": results keep x's strides, the
                # upstream gradient is NCHW, the backward stages three tensors through channels-last copies
                with torch.no_grad():
                    o_probe = c2m_b200.warp_blend(x2, flow, mask, deterministic=det, flags=flags)
                g_like = g2 if o_probe.is_contiguous() else gout
                del o_probe
                other = {"layout": "nchw", "policy": "x converted once in the forward, channels-last results",
                         **time_other(0, g_like)}
                other["gout_nchw"] = time_other(0, g2)
                other["strict"] = time_other(_lib.FLAG_STRICT_LAYOUT, g2)
            else:
                other = {"layout": "nhwc", **time_other(0, g2)}
            del x2, g2

        # ---- secondary figures: the multi-scale sites of the reference, N frames each, all levels
        # of a pyramid run back to back (small levels are launch-bound: reported as they are)
        pyramids = None
        if args.pyramids:
            sets = {
                "generator_encoder": [(32, 256, 512), (64, 128, 256), (128, 64, 128), (256, 32, 64)],
                "motion_decoder": [(64, 64, 128), (128, 32, 64), (256, 16, 32), (512, 8, 16)],
                "image": [(3, 256, 512)],
                # the image warps as the reference differentiates them (generator.py:140 kitti, losses.py:219-222): the
                # frames are data, no mask, only the flow carries a gradient
                "image_flow_grad_only": [(3, 256, 512)],
            }
            pyramids = {}
            for name, levels in sets.items():
                flow_only = name == "image_flow_grad_only"
                ts = []
                for (c, h, w) in levels:
                    px, pf, pm, pg = synth(N, c, h, w, False, 77 + rank, dev)
                    if nhwc and c % 4 == 0:
                        px = px.contiguous(memory_format=torch.channels_last)
                        pg = pg.contiguous(memory_format=torch.channels_last)
                    if flow_only:
                        ts.append((px, pf.requires_grad_(True), None, pg))
                    else:
                        ts.append((px.requires_grad_(True), pf.requires_grad_(True), pm.requires_grad_(True), pg))

                def wrt(lv):
                    return [t for t in lv[:3] if t is not None and t.requires_grad]

                def pstep():
                    for lv in ts:
                        o = c2m_b200.warp_blend(lv[0], lv[1], lv[2])
                        torch.autograd.grad(o, wrt(lv), lv[3])

                pms, _, _ = _time_steps(pstep, 5, 2, barrier)

                # the way a training step runs them: every level's forward, then ONE backward pass over all of them
                # (one hand-off to the autograd engine instead of one per level)
                def pstep_one():
                    outs = [c2m_b200.warp_blend(lv[0], lv[1], lv[2]) for lv in ts]
                    torch.autograd.grad(outs, [t for lv in ts for t in wrt(lv)], [lv[3] for lv in ts])

                oms, _, _ = _time_steps(pstep_one, 5, 2, barrier)
                # the same launches captured once in a CUDA graph and replayed (the entry points only enqueue work
                # on the given stream): the levels without the per-call host time
                gms = None
                try:
                    side = torch.cuda.Stream(dev)
                    side.wait_stream(torch.cuda.current_stream(dev))
                    with torch.cuda.stream(side):
                        pstep()
                    torch.cuda.current_stream(dev).wait_stream(side)
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        pstep()
                    gms, _, _ = _time_steps(graph.replay, 10, 2, barrier)
                    del graph
                except RuntimeError as e:  # capture not possible: keep the eager figure only
                    print(f"bench.py: pyramid graph capture failed: {e}", file=sys.stderr)
                pbytes = sum(fwd_bytes(N, c, h, w) + bwd_bytes(N, c, h, w) for (c, h, w) in levels)
                if flow_only:  # fwd: x, out, flow; bwd: gout, x, flow, grad-flow
                    pbytes = sum(4 * N * h * w * ((2 * c + 2) + (2 * c + 4)) for (c, h, w) in levels)
                pyramids[name] = {"levels": [list(l) for l in levels], "ms": pms,
                                  "achieved": pbytes / (pms * 1e-3) / 1e9, "frac": pbytes / (pms * 1e-3) / 1e9 / peak,
                                  "one_backward": {"ms": oms, "achieved": pbytes / (oms * 1e-3) / 1e9,
                                                   "frac": pbytes / (oms * 1e-3) / 1e9 / peak}}
                if gms:
                    pyramids[name]["cuda_graph"] = {"ms": gms, "achieved": pbytes / (gms * 1e-3) / 1e9,
                                                    "frac": pbytes / (gms * 1e-3) / 1e9 / peak}
                del ts

        # ---- comparative figure: the reference's own GPU path (ops.py:187-202 + generator.py:93 as the unpatched
        # trainer runs it: CPU-built grid copied to the device every call, div / cat / add, grid_sample, multiply)
        torch_cuda = None
        if args.torch_cuda_steps > 0:
            import torch.nn.functional as F

            def ref_step():
                g0 = torch.zeros([N, 2, H, W])
                g0[:, 0] = torch.linspace(-1, 1, W).view(1, 1, W).expand(N, H, W)
                g0[:, 1] = torch.linspace(-1, 1, H).view(1, H, 1).expand(N, H, W)
                g0 = g0.to(dev)
                nf = torch.cat([flow[:, 0:1] / ((W - 1.0) / 2.0), flow[:, 1:2] / ((H - 1.0) / 2.0)], dim=1)
                o = F.grid_sample(x, (g0 + nf).permute(0, 2, 3, 1), mode="bilinear", padding_mode="border",
                                  align_corners=False) * mask
                torch.autograd.grad(o, [x, flow, mask], gout)

            rms, _, _ = _time_steps(ref_step, args.torch_cuda_steps, 1, barrier)
            torch_cuda = {"ms_per_step": rms, "frames_per_s_per_gpu": N / (rms * 1e-3),
                          "what": "torch %s CUDA composition of the reference path on the same tensors "
                                  "(comparison only)" % torch.__version__}

        # ---- end to end through the public API with HOST buffers (pinned), copies inside the timed region
        e2e = None
        if args.e2e_steps > 0:
            hx, hflow, hmask, hgout = synth(N, C, H, W, oob, 1234 + rank, dev, pin=True)
            if nhwc:
                hx = hx.contiguous(memory_format=torch.channels_last).pin_memory()
                hgout = hgout.contiguous(memory_format=torch.channels_last).pin_memory()
            from c2m_b200 import host as chost
            plan = chost.HostWarpPlan(N, C, H, W, dev, chunks=args.e2e_chunks, nhwc=nhwc)
            e2e_steps = max(1, min(args.steps, args.e2e_steps))
            ems, et0, et1 = _time_steps(lambda: plan.run(hx, hflow, hmask, hgout), e2e_steps, 2, barrier)
            e2e_value, _, _ = cdist.aggregate_throughput(N * e2e_steps, ems * e2e_steps, dev)
            plan_bytes_per_frame = plan.h2d_bytes / N
            e2e = {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": plan.h2d_bytes,
                   "d2h_bytes_per_step": plan.d2h_bytes, "steps": e2e_steps, "chunks": plan.chunks,
                   "pcie_gbs_each_way": e2e_value / world * plan.h2d_bytes / N / 1e9, "clocks": clk.window(et0, et1)}
            del plan
            # the ceiling this figure lives under: pinned host <-> device copies of the same tensors, both directions
            # at once on two streams, nothing else -- measured by every rank at the same time (the ranks of one box
            # share the host memory system and the PCIe root complexes)
            dx, dg = torch.empty_like(hx, device=dev), torch.empty_like(hgout, device=dev)
            s_up, s_down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

            def both_ways():
                cur = torch.cuda.current_stream(dev)
                s_up.wait_stream(cur)
                s_down.wait_stream(cur)
                with torch.cuda.stream(s_up):
                    dx.copy_(hx, non_blocking=True)
                with torch.cuda.stream(s_down):
                    hgout.copy_(dg, non_blocking=True)
                cur.wait_stream(s_up)
                cur.wait_stream(s_down)

            cms, _, _ = _time_steps(both_ways, 5, 2, barrier)
            cms = cdist.max_over_ranks(cms, dev)
            ceil_gbs = hx.numel() * 4 / (cms * 1e-3) / 1e9
            e2e["pcie_ceiling_gbs_each_way"] = ceil_gbs
            # per-rank bytes per second each way in the e2e run, over the per-rank ceiling under the same concurrency
            e2e["frac_of_pcie_ceiling"] = (e2e_value / world * plan_bytes_per_frame / 1e9) / ceil_gbs
            del dx, dg, hx, hflow, hmask, hgout

        # ---- the dominant kernels alone: the library brackets its forward kernel / backward gather kernel with
        # CUDA events on the launching stream (c2m_warp_profile); read back after each call, outside the timed region
        kfwd, kbwd = [], []
        _lib.profile(True)
        try:
            for _ in range(5):
                out = c2m_b200.warp_blend(x, flow, mask, deterministic=det, flags=flags)
                kfwd.append(_lib.profile_last_ms())
                torch.autograd.grad(out, [x, flow, mask], gout)
                torch.cuda.synchronize(dev)
                kbwd.append(_lib.profile_last_ms())
            del out
        finally:
            _lib.profile(False)
        kfwd_ms = statistics.median(kfwd) if kfwd and min(kfwd) > 0 else None
        kbwd_ms = statistics.median(kbwd) if kbwd and min(kbwd) > 0 else None

        # ---- BASELINE.json configs[2] (KITTI shape, large / out-of-bounds flows, deterministic on and off) and
        # configs[4] (full-resolution C=256 shards): secondary workloads, each with its own clock window
        configs = None
        if args.configs:
            configs = {}
            configs["fullres_1024x2048_c256"] = measure_config("fullres_1024x2048_c256", dev, rank, world, clk, barrier,
                                                               peak, nhwc=nhwc, steps=5, warmup=3)
            if world == 1:
                configs["kitti_256x832_c64_oob"] = measure_config("kitti_256x832_c64_oob", dev, rank, world, clk,
                                                                  barrier, peak, nhwc=nhwc, steps=10, warmup=3)
                configs["kitti_256x832_c64_oob_deterministic"] = measure_config(
                    "kitti_256x832_c64_oob", dev, rank, world, clk, barrier, peak, nhwc=nhwc, det=True, steps=10, warmup=3)
        blend = None
        if args.configs and world == 1:
            blend = measure_config(workload, dev, rank, world, clk, barrier, peak, nhwc=nhwc, steps=10, warmup=3,
                                   blend=True)

    dominant = "bwd" if bwd_ms >= fwd_ms else "fwd"
    dom_bytes = bwd_bytes(N, C, H, W) if dominant == "bwd" else fwd_bytes(N, C, H, W)
    dom_kernel_ms = kbwd_ms if dominant == "bwd" else kfwd_ms
    dom_ms = dom_kernel_ms or (bwd_ms if dominant == "bwd" else fwd_ms)
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    kname = ("the gather kernel of c2m_warp_blend_bwd (CUDA events around that launch)" if dominant == "bwd"
             else "the forward kernel of c2m_warp_blend_fwd (CUDA events around that launch)")
    if not dom_kernel_ms:
        kname = f"{dominant} (c2m_warp_blend_{dominant}: all launches of the call)"
    fb, bb = fwd_bytes(N, C, H, W), bwd_bytes(N, C, H, W)
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": None, "kernel": kname, "layout": args.layout,
            "peak_source": peak_src, "ms_per_launch": dom_ms,
            "kernels_alone": {"fwd_ms": kfwd_ms, "bwd_gather_ms": kbwd_ms},
            "fwd": {"ms": fwd_ms, "achieved": fb / (fwd_ms * 1e-3) / 1e9, "frac": fb / (fwd_ms * 1e-3) / 1e9 / peak},
            "bwd": {"ms": bwd_ms, "achieved": bb / (bwd_ms * 1e-3) / 1e9, "frac": bb / (bwd_ms * 1e-3) / 1e9 / peak},
            "fwd_bwd": {"ms": fwd_ms + bwd_ms, "achieved": (fb + bb) / ((fwd_ms + bwd_ms) * 1e-3) / 1e9,
                        "frac": (fb + bb) / ((fwd_ms + bwd_ms) * 1e-3) / 1e9 / peak}}
    if other is not None:
        roof["other_layout"] = other
    if pyramids is not None:
        roof["pyramids"] = pyramids
    if configs is not None:
        roof["configs"] = configs
    if blend is not None:
        roof["blend"] = blend
    del x, flow, mask, gout
    torch.cuda.empty_cache()

    cpu = None
    if rank == 0 and world == 1:
        if not args.no_traffic:
            regex = ("gather_nhwc_kernel|gather_nchw_kernel|bwd_scatter_kernel" if dominant == "bwd"
                     else "fwd_nhwc_kernel|fwd_nchw_kernel|fwd_generic_kernel")
            roof["traffic"], roof["traffic_source"] = measure_traffic(args, regex)
        if not args.no_cpu_baseline:
            frames = args.cpu_frames or N
            times, threads = cpu_reference_run(workload, frames, 5, 2)
            med = statistics.median(times)
            cpu = {"value": frames / med, "unit": "frames/s", "cores": threads, "kind": "port",
                   "sample": f"{frames} frames (C={C}, {H}x{W}) fwd+bwd, median of 5 after 2 warm-ups, "
                             f"oracle.reference_torch on torch {torch.__version__} CPU"}
            # BASELINE.json configs[0]: the reference's own CPU-runnable case (forward, 5 frames of 64 x 128 x 256)
            n0, c0, h0, w0, _ = WORKLOADS["cpu_128x256_c64"]
            t0s, _ = cpu_reference_run("cpu_128x256_c64", n0, 10, 3, forward_only=True)
            cpu["configs0_forward"] = {"workload": "cpu_128x256_c64", "value": n0 / statistics.median(t0s),
                                       "unit": "frames/s", "sample": f"{n0} frames (C={c0}, {h0}x{w0}) forward, "
                                                                     "median of 10 after 3 warm-ups"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(workload, N, C, H, W, det), "layout": args.layout,
            "roofline": roof, "cpu_baseline": cpu, "torch_cuda_reference": torch_cuda,
            "e2e": e2e,
            "gpu_launches": launches, "clocks": clk.summary(),
        }
        emit(line)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    torch.cuda.synchronize(dev)
    return 0


class _QuietStdout:
    """Everything any library writes to stdout (fd 1) while the benchmark runs -- NCCL's version banner, for one
    -- is sent to stderr; only the result line goes to the real stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text):
        sys.stdout.flush()
        os.write(self.saved, (text + "\n").encode())

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


_OUT = None


def emit(line):
    text = json.dumps(line)
    if _OUT is not None:
        _OUT.emit(text)
    else:
        print(text, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cityscapes_256x512_c64", choices=sorted(WORKLOADS))
    ap.add_argument("--layout", default="nhwc", choices=["nchw", "nhwc"],
                    help="memory format of x / gout / out / gx (logical shape is always [N,C,H,W])")
    ap.add_argument("--no-other-layout", action="store_true", help="skip the secondary (other layout) measurement")
    ap.add_argument("--frames", type=int, default=0, help="override frames per GPU")
    ap.add_argument("--deterministic", action="store_true")
    ap.add_argument("--flags", default="0")
    ap.add_argument("--cpu-frames", type=int, default=0,
                    help="frames per step of the CPU runs (0: the same as the product arm's step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pyramids", dest="pyramids", action="store_false", help="skip the multi-scale secondary figures")
    ap.add_argument("--no-configs", dest="configs", action="store_false",
                    help="skip the secondary workloads (BASELINE configs[2], configs[4], blend operand)")
    ap.add_argument("--no-traffic", action="store_true", help="skip the ncu child that measures roofline.traffic")
    ap.add_argument("--traffic-child", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--torch-cuda-steps", type=int, default=3, help="steps of the torch CUDA composition (0: skip)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-chunks", type=int, default=8)
    args = ap.parse_args()
    if args.traffic_child:
        return run_traffic_child(args)
    global _OUT
    with _QuietStdout() as q:
        _OUT = q
        try:
            return run_reference(args) if args.impl == "reference" else run_ours(args)
        finally:
            _OUT = None


if __name__ == "__main__":
    # a normal interpreter exit: atexit hooks run (the driver's record of the mapped shared objects among them)
    sys.exit(main())
