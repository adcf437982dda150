"""Head-sharded multi-GPU attention: one process per GPU, every (batch, head) pair is an independent problem
(the reference's grid axes 1, 2 are B and H, flash_attention_torch.py:59), so rank g owns a contiguous head
slice and the compute path has no collective.  The only communication is the optional all-gather (NCCL over
NVLink / NVSwitch via torch.distributed) that reassembles O on every rank when the caller asks for it.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def head_range(H: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous head slice [h0, h1) of rank `rank`; the first H % world ranks take one extra head."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(H, world)
    h0 = rank * base + min(rank, rem)
    return h0, h0 + base + (1 if rank < rem else 0)


def all_gather_heads(local: torch.Tensor, H: int, group=None) -> torch.Tensor:
    """(B, h_local, N, d) on each rank -> (B, H, N, d) on every rank, heads in rank order."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    B, _, N, d = local.shape
    sizes = [head_range(H, r, world) for r in range(world)]
    out = torch.empty((B, H, N, d), dtype=local.dtype, device=local.device)
    if B == 1 and H % world == 0:
        # each rank's slab is contiguous in the result: gather straight into it, no staging
        dist.all_gather_into_tensor(out.view(-1), local.contiguous().view(-1), group=group)
        return out
    # general case (B > 1 or uneven slices): equal-sized staging slabs, then one strided copy per rank
    h_max = max(h1 - h0 for h0, h1 in sizes)
    slab = torch.zeros((B, h_max, N, d), dtype=local.dtype, device=local.device)
    slab[:, : local.shape[1]] = local
    stage = torch.empty((world, B, h_max, N, d), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(stage.view(-1), slab.view(-1), group=group)
    for r, (h0, h1) in enumerate(sizes):
        out[:, h0:h1] = stage[r, :, : h1 - h0]
    assert sizes[rank][1] - sizes[rank][0] == local.shape[1]
    return out


class PeerGatherBuffer:
    """Gathered output (B, H, N, d) that every rank's forward kernel writes DIRECTLY: rank g's epilogue stores its head
    slice into all ranks' copies over NVLink while it is still computing the remaining tiles — through one NVLS
    multicast store per 16 bytes when the switch supports it, otherwise through per-peer P2P stores (fa_fwd_peers).
    Replaces the NCCL all-gather that would follow the kernel.  Built on torch's symmetric memory (allocation,
    handle exchange and barrier are plumbing; the data path is the attention kernel itself)."""

    def __init__(self, B: int, H: int, N: int, d: int, dtype: torch.dtype, device, group=None, use_multicast=True):
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if self.world > 8:
            raise ValueError("peer gather is a single-box path (<= 8 GPUs)")
        self.tensor = symm_mem.empty((B, H, N, d), dtype=dtype, device=device)
        self.handle = symm_mem.rendezvous(self.tensor, self.group)
        self.peer_base = [int(p) for p in self.handle.buffer_ptrs]
        mc = int(getattr(self.handle, "multicast_ptr", 0) or 0)
        self.multicast_base = mc if (use_multicast and mc) else 0
        self.H = H

    def window(self, rank: int):
        """(local view, [(data_ptr, strides)] for the kernel) of rank `rank`'s head slice."""
        h0, h1 = head_range(self.H, rank, self.world)
        return self.tensor[:, h0:h1]

    def forward_into(self, Q_local, K_local, V_local, causal, softmax_scale):
        """Run this rank's forward with the fused gather; returns (gathered tensor, L_local).  The gathered tensor is
        complete on every rank after the barrier this method issues on the current stream.

        Lifetime rule: the returned tensor is the live symmetric buffer — the NEXT call overwrites it on every rank.
        A barrier on entry orders every rank's reads of the previous result (on its current stream) before any rank's
        epilogue stores of this call; work queued on OTHER streams must be ordered by the caller (or use a clone).
        Forward-only: the kernel is called below the autograd Function, the result carries no grad_fn."""
        from . import _native

        if torch.is_grad_enabled() and any(t.requires_grad for t in (Q_local, K_local, V_local)):
            raise RuntimeError("PeerGatherBuffer.forward_into is forward-only (the gathered O has no grad_fn); "
                               "run it under torch.no_grad() or use head_sharded_attention(gather=False) for training")
        view = self.window(self.rank)
        assert view.shape == Q_local.shape, (view.shape, Q_local.shape)
        # write-after-read across calls: nobody may store into a peer's copy while that peer still reads the last result
        self.handle.barrier()
        if Q_local.numel() == 0:   # this rank owns no heads (H < world): nothing to compute, only the barriers
            L = torch.empty(Q_local.shape[:3], dtype=torch.float32, device=Q_local.device)
            self.handle.barrier()
            return self.tensor, L
        off = view.data_ptr() - self.tensor.data_ptr()
        strides = tuple(view.stride())
        if self.multicast_base:
            _, L = _native.forward(Q_local, K_local, V_local, causal, softmax_scale,
                                   out=(self.multicast_base + off, strides))
        else:
            peers = [b + off for r, b in enumerate(self.peer_base) if r != self.rank]
            _, L = _native.forward(Q_local, K_local, V_local, causal, softmax_scale,
                                   out=(view.data_ptr(), strides), peer_ptrs=peers)
        self.handle.barrier()
        return self.tensor, L


class _GatherHeads(torch.autograd.Function):
    """all_gather_heads with a backward: every rank receives the full dO of the gathered output, and the gradient of
    its own head slice is just that slice (each rank's loss is a function of the same gathered O; summing the ranks'
    contributions is the caller's data-parallel reduction, exactly as with any replicated activation)."""

    @staticmethod
    def forward(ctx, local, H, group):
        world = dist.get_world_size(group)
        ctx.h0, ctx.h1 = head_range(H, dist.get_rank(group), world)
        return all_gather_heads(local, H, group)

    @staticmethod
    def backward(ctx, grad_full):
        return grad_full[:, ctx.h0:ctx.h1].contiguous(), None, None


def head_sharded_attention(Q, K, V, causal: bool = False, softmax_scale: float = 1.0, group=None,
                           gather: bool = False, attn_fn=None):
    """Attention over this rank's head slice of replicated (B, H, N, d) inputs.

    Returns the local (B, h_local, N, d) output, or the full (B, H, N, d) output on every rank when
    `gather=True` (NCCL all-gather after the kernel) or `gather=<PeerGatherBuffer>` (the kernel's epilogue writes every
    rank's copy itself, no collective).  The head slice is a strided view (no copy);
    `attn_fn(q, k, v, causal, softmax_scale)` defaults to FlashAttention.apply.

    Gradients: the local output and `gather=True` are differentiable (the all-gather's backward hands each rank the
    dO slice of its own heads, then the local attention backward runs); `gather=<PeerGatherBuffer>` is forward-only
    and raises if an input requires grad while grad mode is on.
    """
    if attn_fn is None:
        from .flash_attention_torch import FlashAttention
        attn_fn = FlashAttention.apply
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    H = Q.shape[1]
    h0, h1 = head_range(H, rank, world)
    if isinstance(gather, PeerGatherBuffer):
        return gather.forward_into(Q[:, h0:h1], K[:, h0:h1], V[:, h0:h1], causal, softmax_scale)[0]
    o_local = attn_fn(Q[:, h0:h1], K[:, h0:h1], V[:, h0:h1], causal, softmax_scale)
    if not gather or world == 1:
        return o_local
    if torch.is_grad_enabled() and o_local.requires_grad:
        return _GatherHeads.apply(o_local, H, group)
    return all_gather_heads(o_local.detach(), H, group)
