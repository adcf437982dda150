"""world_size-2 gloo test of the head-sharded path on CPU: slicing, local compute (the oracle stands in for the
CUDA kernel, which cannot run here) and the optional all-gather reassembly."""
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


def _worker(rank, world, port, B, H, ok):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from flash_attention_dlrs_b200.sharding import head_range, head_sharded_attention
        from oracle import attention_oracle as orc

        torch.manual_seed(11)  # replicated inputs
        Q, K, V = (torch.randn(B, H, 48, 16) for _ in range(3))
        fn = lambda q, k, v, c, s: orc.reference_sdpa(q, k, v, s, c)
        full = orc.reference_sdpa(Q, K, V, 0.25, True)
        local = head_sharded_attention(Q, K, V, True, 0.25, gather=False, attn_fn=fn)
        h0, h1 = head_range(H, rank, world)
        good = local.shape[1] == h1 - h0 and torch.allclose(local, full[:, h0:h1], atol=1e-6)
        gathered = head_sharded_attention(Q, K, V, True, 0.25, gather=True, attn_fn=fn)
        good = good and gathered.shape == full.shape and torch.allclose(gathered, full, atol=1e-6)
        # training through the gathered output: every rank gets the gradients of ITS head slice of the replicated inputs
        Qg, Kg, Vg = (t.clone().requires_grad_(True) for t in (Q, K, V))
        w = torch.randn(full.shape, generator=torch.Generator().manual_seed(5))
        out = head_sharded_attention(Qg, Kg, Vg, True, 0.25, gather=True, attn_fn=fn)
        good = good and out.requires_grad
        (out * w).sum().backward()
        Qr, Kr, Vr = (t.clone().requires_grad_(True) for t in (Q, K, V))
        (orc.reference_sdpa(Qr, Kr, Vr, 0.25, True) * w).sum().backward()
        for got, ref in ((Qg.grad, Qr.grad), (Kg.grad, Kr.grad), (Vg.grad, Vr.grad)):
            good = good and torch.allclose(got[:, h0:h1], ref[:, h0:h1], atol=1e-5)
            rest = torch.cat([got[:, :h0], got[:, h1:]], dim=1)
            good = good and not rest.any()          # other ranks' heads: no gradient on this rank
        ok[rank] = 1 if good else 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B,H", [(1, 4), (2, 5)])
def test_head_sharding_world2_gloo(B, H):
    world = 2
    ok = mp.get_context("spawn").Array("i", [0] * world)
    mp.spawn(_worker, args=(world, _free_port(), B, H, ok), nprocs=world, join=True)
    assert list(ok) == [1] * world
