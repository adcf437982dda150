#!/usr/bin/env python
"""Sequence-parallel (ring) attention on the GPUs of one box (torchrun): rank r owns sequence shard r of a (B, H, N, D)
problem; checks O / L / dQ / dK / dV of its shard against full attention computed locally and times ring forward+backward
next to the single-GPU kernels on the whole sequence.   args: B H N D"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from flash_attention_dlrs_b200 import _native, ring

B, H, N, D = 1, 8, 32768, 128
if len(sys.argv) > 4:
    B, H, N, D = (int(x) for x in sys.argv[1:5])
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = N // world
scale = D ** -0.5
g = torch.Generator(device=dev).manual_seed(7)          # same full tensors on every rank
Q, K, V, dO = (torch.randn(B, H, N, D, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16) for _ in range(4))


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item(), out


for causal, zigzag, peer in ((False, False, False), (False, False, True), (True, False, True), (True, True, False),
                             (True, True, True)):
    shard = (lambda t: ring.zigzag_shard(t, rank, world).contiguous()) if zigzag else \
        (lambda t: t[:, :, rank * n:(rank + 1) * n].contiguous())
    q, k, v, do = (shard(t) for t in (Q, K, V, dO))

    tr = ring.PeerTransport(B, H, n, D, torch.bfloat16, dev, zigzag=zigzag) if peer else None

    def ring_step():
        O, L = ring.ring_attention_forward(q, k, v, causal, scale, zigzag=zigzag, transport=tr)
        return (O, L) + ring.ring_attention_backward(q, k, v, O, do, L, causal, scale, zigzag=zigzag, transport=tr)

    def full_step():
        O, L = _native.forward(Q, K, V, causal, scale)
        return (O, L) + _native.backward(Q, K, V, O, dO, L, causal, scale)

    ms_ring, got = timed(ring_step)
    ms_full, want = timed(full_step)
    rel = lambda a, b: ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-6)).item()
    errs = {"O": (got[0].float() - shard(want[0]).float()).abs().max().item(),
            "L": (got[1] - shard(want[1].unsqueeze(-1))).abs().max().item(),
            "dQ": rel(got[2], shard(want[2])), "dK": rel(got[3], shard(want[3])), "dV": rel(got[4], shard(want[4]))}
    ok = errs["O"] <= 2e-2 and errs["L"] <= 2e-3 and max(errs["dQ"], errs["dK"], errs["dV"]) <= 2e-2
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        fl = 3.5 * 4.0 * B * H * N * N * D * (0.5 if causal else 1.0)
        print(json.dumps({"config": f"ring fwd+bwd bf16 B={B} H={H} N={N} D={D} causal={causal} zigzag={zigzag} "
                                    f"transport={'peer-memory' if peer else 'nccl-p2p'}", "n_gpus": world,
                          "ok": bool(flag.item()), "errors_rank0_vs_single_gpu_kernels": errs, "ring_ms": ms_ring,
                          "ring_tflops": fl / ms_ring / 1e9, "single_gpu_full_sequence_ms": ms_full}), flush=True)
dist.destroy_process_group()
