#!/usr/bin/env python
"""Head-sharded forward with the FUSED all-gather epilogue (PeerGatherBuffer: NVLS multicast or P2P stores from the
kernel's epilogue) against forward + NCCL all-gather, on the GPUs of one box (torchrun).  Checks that both give the same
gathered O bit for bit and times them (config 4 shape by default: bf16 B=1 H=64 N=32768 D=128 causal)."""
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
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
h0, h1 = sharding.head_range(H, rank, world)
scale = D ** -0.5
g = torch.Generator(device=dev).manual_seed(100 + rank)
Q, K, V = (torch.randn(B, h1 - h0, N, D, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16) for _ in range(3))


def timed(fn, reps=5):
    for _ in range(2):
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


def nccl_path():
    O, L = _native.forward(Q, K, V, True, scale)
    return sharding.all_gather_heads(O, H)


def kernel_only():
    return _native.forward(Q, K, V, True, scale)[0]


res = {"config": f"fwd bf16 B={B} H={H} N={N} D={D} causal, head-sharded, O gathered on every rank", "n_gpus": world}
res["fwd_only_ms"], _ = timed(kernel_only)
res["fwd_plus_nccl_allgather_ms"], want = timed(nccl_path)
for mc in (True, False):
    try:
        buf = sharding.PeerGatherBuffer(B, H, N, D, torch.bfloat16, dev, use_multicast=mc)
        mode = "multicast" if buf.multicast_base else "p2p"
        if not mc and mode != "p2p":
            continue
        if mc and mode == "p2p":
            res["multicast"] = "not supported here"
            continue
        ms, got = timed(lambda: buf.forward_into(Q, K, V, True, scale)[0])
        res[f"fwd_fused_gather_{mode}_ms"] = ms
        res[f"fused_{mode}_equals_nccl_bitwise"] = bool(torch.equal(got, want))
    except Exception as e:   # noqa: BLE001
        res[f"fused_gather_error_mc{int(mc)}"] = repr(e)[:300]
if rank == 0:
    print(json.dumps(res), flush=True)
dist.destroy_process_group()
