#!/usr/bin/env python
"""BASELINE config 4 — long-context fwd+bwd bf16 B=1 H=64 N=32768 D=128 causal, head-sharded over the GPUs of one box
(strong scaling: the job is fixed, rank g owns heads [g*H/G, (g+1)*H/G)).  Launch with torchrun (or plain python for G=1):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P tools/multi_gpu_c4.py

No collective on the compute path; the optional NCCL all-gather of O (what a caller that wants O on one device pays) is
timed separately.  Every head's inputs are seeded by the head index, so the per-head output fingerprints printed by
rank 0 must be identical for every G (sharding does not change a single bit).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from flash_attention_dlrs_b200 import _native, sharding

B, H, N, D = 1, 64, 32768, 128
if len(sys.argv) > 4:
    B, H, N, D = (int(x) for x in sys.argv[1:5])
steps, warmup = 5, 2
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
h0, h1 = sharding.head_range(H, rank, world)
hl = h1 - h0
scale = D ** -0.5


def make(kind):
    out = torch.empty(B, hl, N, D, dtype=torch.bfloat16, device=dev)
    for i, h in enumerate(range(h0, h1)):
        g = torch.Generator(device=dev).manual_seed(1000 * h + kind)
        out[:, i] = torch.randn(B, N, D, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    return out


Q, K, V, dO = (make(k) for k in range(4))


def step():
    O, L = _native.forward(Q, K, V, True, scale)
    return O, _native.backward(Q, K, V, O, dO, L, True, scale)


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for _ in range(warmup):
    step()
sync()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(steps):
    O, (dQ, dK, dV) = step()
b.record()
sync()
ms = torch.tensor([a.elapsed_time(b) / steps], device=dev)
gather_ms = torch.zeros(1, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    full = sharding.all_gather_heads(O, H)          # warm-up
    sync()
    a.record()
    full = sharding.all_gather_heads(O, H)
    b.record()
    sync()
    gather_ms[0] = a.elapsed_time(b)
    dist.all_reduce(gather_ms, op=dist.ReduceOp.MAX)
# per-head fingerprints of O and dQ (exact integer sums of the raw 16-bit patterns)
fp = torch.stack([torch.stack([t[:, i].contiguous().view(torch.int16).to(torch.int64).sum() for t in (O, dQ, dK, dV)])
                  for i in range(hl)])
if world > 1:
    parts = [torch.zeros_like(fp) for _ in range(world)] if H % world == 0 else None
    dist.all_gather(parts, fp)
    fp = torch.cat(parts)
if rank == 0:
    fl = 3.5 * 4.0 * B * H * N * N * D * 0.5
    import hashlib
    digest = hashlib.sha256(fp.cpu().numpy().tobytes()).hexdigest()[:16]
    print(json.dumps({"config": f"C4 fwd+bwd bf16 B={B} H={H} N={N} D={D} causal, head-sharded", "n_gpus": world,
                      "ms_per_step": ms.item(), "tflops": fl / (ms.item() * 1e-3) / 1e12, "scaling": "strong",
                      "allgather_O_ms": gather_ms.item(), "allgather_bytes_per_rank": B * hl * N * D * 2,
                      "fingerprint_O_dQ_dK_dV": digest}), flush=True)
if world > 1:
    dist.destroy_process_group()
