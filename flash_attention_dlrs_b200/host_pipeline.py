"""Attention for callers whose tensors live in HOST memory (pinned): forward + backward with the host<->device copies
pipelined against the kernels.

Every (batch, head) pair is an independent problem, so the (B*H) axis is cut into chunks; chunk c+1 is copied in on
a copy stream while chunk c runs on the compute stream and the results of chunk c-1 are copied out on a third
stream.  PCIe is full duplex, so the step costs about max(H2D, D2H, compute) instead of their sum.  Consecutive
run() calls pipeline into each other as well: the staging buffers are guarded by events that persist across calls, so
the first uploads of step k+1 overlap the last kernels and downloads of step k (no fill / drain bubble per step).
The kernels are the same C-ABI entry points (`_native.forward` / `_native.backward`); nothing is computed on the host.
"""
from __future__ import annotations

import torch

from . import _native


def _as_bh(t: torch.Tensor, what: str) -> torch.Tensor:
    """(B, H, N, D) -> (B*H, N, D) as a VIEW: a reshape that copies would send the device->host results into a
    temporary (outputs) or add a hidden pageable copy (inputs), so non-contiguous or unpinned buffers are refused."""
    if t.device.type != "cpu" or not t.is_pinned() or not t.is_contiguous():
        raise ValueError(f"{what} must be contiguous pinned host tensors of shape (B, H, N, D)")
    B, H, N, D = t.shape
    return t.view(B * H, N, D)


class HostAttentionPipeline:
    """Reusable pipeline for one problem shape: owns the device staging buffers, streams and events."""

    def __init__(self, B: int, H: int, N: int, D: int, dtype: torch.dtype, device, chunks: int = 8,
                 with_backward: bool = True, duplex: bool = True):
        """duplex=True puts host->device and device->host copies on separate streams (full-duplex PCIe); on hosts
        where simultaneous traffic in both directions collapses the link rate, duplex=False keeps one copy stream and
        only overlaps the copies with the kernels."""
        self.shape = (B, H, N, D)
        self.dtype, self.device = dtype, torch.device(device)
        BH = B * H
        chunks = max(1, min(chunks, BH))
        while BH % chunks:
            chunks -= 1
        self.chunks, self.per = chunks, BH // chunks
        self.with_backward = with_backward
        n_in = 4 if with_backward else 3
        n_out = 4 if with_backward else 1
        mk = lambda: torch.empty((1, self.per, N, D), dtype=dtype, device=self.device)
        # double-buffered inputs, double-buffered outputs
        self.inp = [[mk() for _ in range(n_in)] for _ in range(2)]
        self.out = [[mk() for _ in range(n_out)] for _ in range(2)]
        self.lse = [torch.empty((1, self.per, N), dtype=torch.float32, device=self.device) for _ in range(2)]
        self.s_in = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device) if duplex else self.s_in
        self.ev_in = [torch.cuda.Event() for _ in range(2)]
        self.ev_done = [torch.cuda.Event() for _ in range(2)]
        self.ev_in_free = [torch.cuda.Event() for _ in range(2)]
        self.ev_out_free = [torch.cuda.Event() for _ in range(2)]

    def run(self, host_in, host_out, causal: bool = False, softmax_scale: float = 1.0, lse_out=None):
        """host_in = (Q, K, V[, dO]) pinned (B,H,N,D) tensors; host_out = (O[, dQ, dK, dV]) pinned tensors written in
        place; `lse_out` (optional): pinned float32 (B,H,N) or (B,H,N,1) that receives L (log2 units).  Asynchronous with respect to the host, and the calling stream does not wait for the copies either: the
        caller synchronises on the returned event (recorded after the last device->host copy) before reading
        `host_out` or rewriting `host_in`.  Calls must come from one stream (the staging buffers are ordered by events
        recorded on it)."""
        compute = torch.cuda.current_stream(self.device)
        hin = [_as_bh(t, "host_in") for t in host_in]
        hout = [_as_bh(t, "host_out") for t in host_out]
        if lse_out is not None:
            lse_v = _as_bh(lse_out.view(*lse_out.shape[:3], 1), "lse_out")[..., 0]   # (B*H, N) float32
        per, nc = self.per, self.chunks
        # ev_in_free / ev_out_free carry over from the previous call (an event never recorded does not block)

        def copy_in(c):
            b = c & 1
            with torch.cuda.stream(self.s_in):
                self.s_in.wait_event(self.ev_in_free[b])
                for dst, src in zip(self.inp[b], hin):
                    dst[0].copy_(src[c * per:(c + 1) * per], non_blocking=True)
                self.ev_in[b].record(self.s_in)

        copy_in(0)
        for c in range(nc):
            b = c & 1
            if c + 1 < nc:
                copy_in(c + 1)
            compute.wait_event(self.ev_in[b])
            compute.wait_event(self.ev_out_free[b])
            q, k, v = self.inp[b][:3]
            O, L = _native.forward(q, k, v, causal, softmax_scale)
            outs = [O]
            if self.with_backward:
                outs += list(_native.backward(q, k, v, O, self.inp[b][3], L, causal, softmax_scale))
            for dst, src in zip(self.out[b], outs):
                dst.copy_(src)          # device-side staging so the allocator can recycle `outs` immediately
            if lse_out is not None:
                self.lse[b].copy_(L)
            self.ev_in_free[b].record(compute)
            self.ev_done[b].record(compute)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_done[b])
                for dst, src in zip(hout, self.out[b]):
                    dst[c * per:(c + 1) * per].copy_(src[0], non_blocking=True)
                if lse_out is not None:
                    lse_v[c * per:(c + 1) * per].copy_(self.lse[b][0], non_blocking=True)
                self.ev_out_free[b].record(self.s_out)
        done = torch.cuda.Event()
        done.record(self.s_out)
        return done


def attention_from_host(Q, K, V, dO=None, causal: bool = False, softmax_scale: float = 1.0, device="cuda",
                        chunks: int = 8, out=None, duplex: bool = True):
    """One-shot convenience wrapper: pinned host tensors in, pinned host tensors out (O, or O, dQ, dK, dV with dO)."""
    if Q.dim() != 4 or Q.shape != K.shape or Q.shape != V.shape:
        raise ValueError("Q, K, V must all be of shape (B, H, N, d)")
    if Q.device.type != "cpu" or not Q.is_pinned():
        raise ValueError("attention_from_host expects pinned host tensors")
    _native.dtype_code(Q.dtype)
    B, H, N, D = Q.shape
    pipe = HostAttentionPipeline(B, H, N, D, Q.dtype, device, chunks, with_backward=dO is not None, duplex=duplex)
    n_out = 4 if dO is not None else 1
    if out is None:
        out = [torch.empty(Q.shape, dtype=Q.dtype).pin_memory() for _ in range(n_out)]
    ins = (Q, K, V) if dO is None else (Q, K, V, dO)
    pipe.run(ins, out, causal, softmax_scale).synchronize()
    return out[0] if dO is None else tuple(out)
