"""Host-buffer entry point: the fused warp forward+backward called with HOST (pinned) tensors, the
way a plugin embedded in a CPU-side pipeline would call it.

The batch x frame axis is cut into chunks that flow through three CUDA streams -- host-to-device
copy, compute (C ABI on device pointers), device-to-host copy -- so the two PCIe directions and the
kernels overlap; nothing is computed on the host.
"""
from __future__ import annotations

import torch

from . import _lib
from .dist import shard_range


class HostWarpPlan:
    """Pre-allocated device staging + pinned host outputs for N frames of [C,H,W].

    run(hx, hflow, hmask, hgout) -> (out, gx, gflow, gmask) pinned host tensors, valid once the
    current stream has been synchronised (the call itself is asynchronous).  The returned tensors are
    the plan's own buffers: the next run() overwrites them, and consecutive calls overlap (the next
    call's uploads start while this call's results are still going out).
    """

    def __init__(self, N, C, H, W, device, chunks=8, nhwc=False, padding="border", deterministic=False):
        self.N, self.C, self.H, self.W = N, C, H, W
        self.device = torch.device(device)
        self.chunks = max(1, min(chunks, N))
        self.padding = _lib.PAD_BORDER if padding == "border" else _lib.PAD_ZEROS
        self.flags = _lib.FLAG_DETERMINISTIC if deterministic else 0
        fmt = torch.channels_last if nhwc else torch.contiguous_format
        d = self.device
        big = dict(dtype=torch.float32, device=d, memory_format=fmt)
        self.x = torch.empty((N, C, H, W), **big)
        self.gout = torch.empty((N, C, H, W), **big)
        self.out = torch.empty((N, C, H, W), **big)
        self.gx = torch.empty((N, C, H, W), **big)
        self.flow = torch.empty((N, 2, H, W), dtype=torch.float32, device=d)
        self.mask = torch.empty((N, 1, H, W), dtype=torch.float32, device=d)
        self.gflow = torch.empty_like(self.flow)
        self.gmask = torch.empty_like(self.mask)
        pin = dict(dtype=torch.float32, pin_memory=True)
        self.h_out = torch.empty((N, C, H, W), memory_format=fmt, **pin)
        self.h_gx = torch.empty((N, C, H, W), memory_format=fmt, **pin)
        self.h_gflow = torch.empty((N, 2, H, W), **pin)
        self.h_gmask = torch.empty((N, 1, H, W), **pin)
        # chunk sizes differ by at most one frame; the workspace is not monotonic in the frame count (small
        # chunks of many-channel levels are channel-sliced and need partial-sum buffers), so size for each
        sizes = {b - a for a, b in (shard_range(N, k, self.chunks) for k in range(self.chunks)) if b > a}
        ws_bytes = max(_lib.bwd_workspace_bytes(n, C, H, W, n, True, self.flags) for n in sizes)
        self.ws = [torch.empty(ws_bytes, dtype=torch.uint8, device=d) for _ in range(2)]
        self.s_in = torch.cuda.Stream(device=d)
        self.s_cmp = torch.cuda.Stream(device=d)
        self.s_out = torch.cuda.Stream(device=d)
        # per chunk: "compute has read the staged inputs" / "the results have left the device" of the previous
        # run() -- consecutive calls pipeline into each other instead of draining the three streams in between
        self.ev_cmp = [None] * self.chunks
        self.ev_out = [None] * self.chunks
        self.primed = False
        elems_in = N * H * W * (2 * C + 3)
        elems_out = N * H * W * (2 * C + 3)
        self.h2d_bytes = 4 * elems_in
        self.d2h_bytes = 4 * elems_out

    def run(self, hx, hflow, hmask, hgout):
        N, C, H, W = self.N, self.C, self.H, self.W
        cur = torch.cuda.current_stream(self.device)
        if not self.primed:  # first call: whatever allocated / touched the staging buffers on `cur` comes first
            for s in (self.s_in, self.s_cmp, self.s_out):
                s.wait_stream(cur)
            self.primed = True
        with torch.cuda.device(self.device):
            for k in range(self.chunks):
                a, b = shard_range(N, k, self.chunks)
                if a == b:
                    continue
                sl = slice(a, b)
                # staging buffers are reused by the next call: chunk k's inputs may be overwritten once the previous
                # call's kernels have read them, its results once they have been copied out
                if self.ev_cmp[k] is not None:
                    self.s_in.wait_event(self.ev_cmp[k])
                if self.ev_out[k] is not None:
                    self.s_cmp.wait_event(self.ev_out[k])
                with torch.cuda.stream(self.s_in):
                    self.x[sl].copy_(hx[sl], non_blocking=True)
                    self.flow[sl].copy_(hflow[sl], non_blocking=True)
                    self.mask[sl].copy_(hmask[sl], non_blocking=True)
                    self.gout[sl].copy_(hgout[sl], non_blocking=True)
                    ev_in = self.s_in.record_event()
                self.s_cmp.wait_event(ev_in)
                ws = self.ws[k & 1]
                xs, os_ = self.x[sl].stride(), self.out[sl].stride()
                st = self.s_cmp.cuda_stream
                _lib.warp_blend_fwd(self.x[sl].data_ptr(), self.flow[sl].data_ptr(), self.mask[sl].data_ptr(), None,
                                    self.out[sl].data_ptr(), b - a, C, H, W, b - a, xs, os_, self.padding,
                                    self.flags, st)
                _lib.warp_blend_bwd(self.x[sl].data_ptr(), self.flow[sl].data_ptr(), self.mask[sl].data_ptr(), None,
                                    self.gout[sl].data_ptr(), self.gx[sl].data_ptr(), self.gflow[sl].data_ptr(),
                                    self.gmask[sl].data_ptr(), None, b - a, C, H, W, b - a, xs,
                                    self.gout[sl].stride(), self.padding, self.flags, ws.data_ptr(), ws.numel(), st)
                ev_c = self.s_cmp.record_event()
                self.ev_cmp[k] = ev_c
                self.s_out.wait_event(ev_c)
                with torch.cuda.stream(self.s_out):
                    self.h_out[sl].copy_(self.out[sl], non_blocking=True)
                    self.h_gx[sl].copy_(self.gx[sl], non_blocking=True)
                    self.h_gflow[sl].copy_(self.gflow[sl], non_blocking=True)
                    self.h_gmask[sl].copy_(self.gmask[sl], non_blocking=True)
                    self.ev_out[k] = self.s_out.record_event()
        # the caller's stream sees the results of THIS call (the plan's own streams run ahead into the next one)
        cur.wait_stream(self.s_out)
        return self.h_out, self.h_gx, self.h_gflow, self.h_gmask
