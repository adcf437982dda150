"""Ring (sequence-parallel) attention on CPU over gloo: the product schedule of flash_attention_dlrs_b200/ring.py —
K / V shards travelling round the ring, log2-domain merge of the partials, dK / dV accumulators travelling home — with the
oracle supplying each per-shard attention partial (the CUDA kernels cannot run here).  The assembled result must equal
full attention over the concatenated sequence."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _partials():
    from oracle import attention_oracle as orc

    def fwd(q, k, v, causal, scale):
        O, L = orc.attention_fp64(q, k, v, scale, causal)
        return O.float(), L.squeeze(-1).float()

    def bwd(q, k, v, o, do, L, causal, scale, delta):
        q, k, v, do = (t.double() for t in (q, k, v, do))
        S = scale * (q @ k.transpose(-1, -2))
        if causal:
            n = q.shape[-2]
            S = S.masked_fill(~torch.ones(n, n, dtype=torch.bool).tril(), float("-inf"))
        P = torch.exp2(S * orc.LOG2_E - L.double().unsqueeze(-1))            # flash_attention_kernels.py:285
        dV = P.transpose(-1, -2) @ do
        dS = P * (do @ v.transpose(-1, -2) - delta.double().unsqueeze(-1))   # :289-291
        return (scale * dS @ k).float(), (scale * dS.transpose(-1, -2) @ q).float(), dV.float()

    return fwd, bwd


def _worker(rank, world, port, causal, zigzag, ok):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from flash_attention_dlrs_b200 import ring
        from oracle import attention_oracle as orc

        torch.manual_seed(3)  # replicated full tensors; every rank slices its own shard
        B, H, n, d, scale = 2, 3, 24, 16, 0.3
        Q, K, V, dO = (torch.randn(B, H, n * world, d) for _ in range(4))
        ref = orc.attention_grads_fp64(Q, K, V, dO, scale, causal)
        if zigzag:
            shard = lambda t: ring.zigzag_shard(t, rank, world)
        else:
            shard = lambda t: t[:, :, rank * n:(rank + 1) * n]
        ops = ring.TorchOps(*_partials())
        q, k, v, do = (shard(t) for t in (Q, K, V, dO))
        O, L = ring.ring_attention_forward(q, k, v, causal, scale, ops=ops, zigzag=zigzag)
        dQ, dK, dV = ring.ring_attention_backward(q, k, v, O, do, L, causal, scale, ops=ops, zigzag=zigzag)
        good = True
        for got, key in ((O, "O"), (L, "L"), (dQ, "dQ"), (dK, "dK"), (dV, "dV")):
            good = good and torch.allclose(got.double(), shard(ref[key]), atol=2e-5, rtol=1e-5)
        ok[rank] = 1 if good else 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,causal,zigzag", [(2, False, False), (2, True, False), (3, True, False), (2, True, True),
                                                 (3, True, True), (2, False, True)])
def test_ring_attention_schedule_gloo(world, causal, zigzag):
    ok = mp.get_context("spawn").Array("i", [0] * world)
    mp.spawn(_worker, args=(world, _free_port(), causal, zigzag, ok), nprocs=world, join=True)
    assert list(ok) == [1] * world
