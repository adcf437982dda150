"""Sequence-parallel ("ring") attention over the GPUs of one box — SURVEY.md §8f-5.  The reference is single-GPU; this is
the multi-GPU path for a sequence that does not fit (or does not have enough heads to shard by head, sharding.py).

Rank r of G holds the contiguous sequence shard r of Q, K, V: tensors (B, H, n, D), tokens [r*n, (r+1)*n).  The K / V
shards travel round the ring (NCCL point-to-point over NVLink, posted one step ahead so the transfer overlaps the
kernels); at every step the rank runs the ordinary single-GPU kernels on (its Q shard, the visiting K / V shard) and folds
the result in:

  forward   (O_s, L_s) = fa_fwd(Q_r, K_c, V_c)  ->  fa_merge_partial: L = log2(2^L + 2^L_s), O = O 2^(L_old-L) + O_s 2^(L_s-L)
            — the forward kernel's own log2-domain online-softmax bookkeeping (flash_attention_kernels.py:93-97,105-106)
            one level up; the running O stays in fp32 until the end.
  backward  (dQ_s, dK_s, dV_s) = fa_bwd(Q_r, K_c, V_c, dO_r, L_r, delta_r) with the GLOBAL L_r and delta_r = rowsum(dO_r o O_r);
            dQ accumulates locally, the fp32 dK / dV accumulators travel with their K / V shard and arrive home after
            G hops.  Every accumulator has one writer per step and a fixed order of additions: deterministic.

Causal masking (keep key <= query, top-left aligned like flash_attention_openai_tutorial.py:50): a visiting chunk is
skipped when it lies after the query chunk, runs the causal kernel when it is the same chunk and the unmasked kernel
when it lies before.  Shards must have equal length.  With plain sharding (chunk r on rank r) the last rank does G times
the work of the first under the causal mask; `zigzag=True` gives every rank chunks r and 2G-1-r of a sequence cut into
2G chunks, and then every rank runs the same number of block products at every step.

The kernels and the merge are CUDA (libfa_b200.so); `ops` lets the CPU tests run the same schedule with torch arithmetic
and oracle partials over gloo.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib, _native


# ------------------------------------------------------------------------------------------------ arithmetic back ends
class CudaOps:
    """The product path: libfa_b200.so kernels on torch's current stream."""

    @staticmethod
    def fwd(q, k, v, causal, scale):
        if q.shape[2] != k.shape[2]:      # rectangular partial (fa_fwd_rect; never causal)
            return _native.forward_rect(q, k, v, scale)
        return _native.forward(q, k, v, causal, scale)            # O (B,H,n,D) 16-bit, L (B,H,n) fp32 (log2 units)

    @staticmethod
    def bwd(q, k, v, o, do, L, causal, scale, delta):
        if q.shape[2] != k.shape[2]:
            return _native.backward_rect(q, k, v, o, do, L, scale, delta=delta)
        return _native.backward(q, k, v, o, do, L, causal, scale, delta=delta)

    @staticmethod
    def delta(o, do):
        return _native.backward_preprocess(o, do)

    @staticmethod
    def _call(fn, *args):
        _lib.check(fn(*args), fn.__name__)

    @classmethod
    def merge(cls, o_acc, l_acc, o_part, l_part, first):
        lib = _lib.load()
        o_part, l_part = o_part.contiguous(), l_part.contiguous()
        cls._call(lib.fa_merge_partial, _native._ptr(o_acc), _native._ptr(l_acc), _native._ptr(o_part), _native._ptr(l_part),
                  o_acc.numel() // o_acc.shape[-1], o_acc.shape[-1], _native.dtype_code(o_part.dtype), int(first),
                  _native._stream_ptr(o_acc.device))

    @classmethod
    def accumulate(cls, acc, part, first):
        lib = _lib.load()
        part = part.contiguous()
        cls._call(lib.fa_accumulate, _native._ptr(acc), _native._ptr(part), acc.numel(), _native.dtype_code(part.dtype),
                  int(first), _native._stream_ptr(acc.device))

    @classmethod
    def round(cls, acc, dtype):
        lib = _lib.load()
        out = torch.empty(acc.shape, dtype=dtype, device=acc.device)
        cls._call(lib.fa_round_rows, _native._ptr(out), _native._ptr(acc), acc.numel(), _native.dtype_code(dtype),
                  _native._stream_ptr(acc.device))
        return out


class TorchOps:
    """Checker back end for the CPU tests: the same schedule with torch arithmetic; the per-shard attention partials come
    from functions the test injects (the oracle).  Never used by the product path."""

    def __init__(self, fwd, bwd):
        self.fwd, self.bwd = fwd, bwd

    @staticmethod
    def delta(o, do):
        return (o.double() * do.double()).sum(-1).float()

    @staticmethod
    def merge(o_acc, l_acc, o_part, l_part, first):
        if first:
            o_acc.copy_(o_part.float()), l_acc.copy_(l_part)
            return
        l_new = torch.logaddexp2(l_acc, l_part)
        wa = torch.where(torch.isinf(l_acc), torch.zeros_like(l_acc), torch.exp2(l_acc - l_new))
        wp = torch.where(torch.isinf(l_part), torch.zeros_like(l_part), torch.exp2(l_part - l_new))
        o_acc.mul_(wa.unsqueeze(-1)).add_(o_part.float() * wp.unsqueeze(-1))
        l_acc.copy_(l_new)

    @staticmethod
    def accumulate(acc, part, first):
        acc.copy_(part.float()) if first else acc.add_(part.float())

    @staticmethod
    def round(acc, dtype):
        return acc.to(dtype)


# ------------------------------------------------------------------------------------------------ ring plumbing
class _Ring:
    """Rotation of buffers round the ring: post(buffers) starts sending them to rank+1 and receiving rank-1's into fresh
    storage; wait() returns what arrived."""

    def __init__(self, group):
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.next = dist.get_global_rank(group, (self.rank + 1) % self.world) if group is not None else (self.rank + 1) % self.world
        self.prev = dist.get_global_rank(group, (self.rank - 1) % self.world) if group is not None else (self.rank - 1) % self.world
        self.reqs, self.recv = [], []

    def post(self, buffers, into=None):
        """Send `buffers` to rank+1 and receive rank-1's into `into` (caller-owned storage that is reused from step to
        step: fresh allocations of this size inside the loop cost milliseconds of allocator and NCCL buffer churn)."""
        self.recv = into if into is not None else [torch.empty_like(b) for b in buffers]
        ops = []
        for b, r in zip(buffers, self.recv):
            ops.append(dist.P2POp(dist.isend, b if b.is_contiguous() else b.contiguous(), self.next, group=self.group))
            ops.append(dist.P2POp(dist.irecv, r, self.prev, group=self.group))
        self.reqs = dist.batch_isend_irecv(ops) if ops else []

    def wait(self):
        for q in self.reqs:
            q.wait()
        self.reqs = []
        return self.recv


class NcclTransport:
    """K / V shards and dK / dV accumulators move with torch.distributed point-to-point operations (NCCL on GPUs, gloo
    in the CPU tests), posted one step ahead of their use."""

    def __init__(self, group=None):
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def begin(self, K, V, chunk_shapes=None):
        self.kv_ring, self.acc_ring = _Ring(self.group), _Ring(self.group)
        G = self.world
        self.kv_bufs = [[torch.empty_like(K, memory_format=torch.contiguous_format) for _ in range(2)]
                        for _ in range(min(G - 1, 2))]
        self.kv = [K, V]
        if chunk_shapes is not None:
            mk = lambda shp: torch.empty(shp, dtype=torch.float32, device=K.device)
            self.acc_bufs = [[mk(shp) for shp in chunk_shapes for _ in range(2)] for _ in range(2)]
            self.acc = self.acc_bufs[0]
        return self.kv

    def prefetch_kv(self, s):
        if s + 1 < self.world:
            self.kv_ring.post(self.kv, self.kv_bufs[s & 1])

    def next_kv(self, s):
        self.kv = self.kv_ring.wait()
        return self.kv

    def first_acc(self):
        return self.acc

    def recv_acc(self, s):
        self.acc = self.acc_ring.wait()
        return self.acc

    def send_acc(self, s):
        self.acc_ring.post(self.acc, self.acc_bufs[(s + 1) & 1])

    def final_acc(self):
        return self.acc_ring.wait()


class PeerTransport:
    """NVLink peer-memory transport: every rank keeps two K / V slots and two accumulator slots in torch symmetric memory
    and PULLS what it needs from rank-1's slot with a plain device copy: 650 GB/s per direction measured on B200 against
    77 GB/s for NCCL send/recv (which, besides, needs SMs that the attention CTAs occupy).  The pulls run on a SIDE
    stream one step ahead of their use, behind a device-side barrier on a channel of their own, so that neither the
    ~10 us barrier nor the copy sits between two steps' kernels:
      K / V of step s+1 are pulled while step s computes (forward and backward);
      the dK / dV sums of step s are pulled while step s's backward kernels run, and only the (memory-bound) additions
      wait for them.
    Build it once per problem shape (the rendezvous is a collective) and pass it as `transport=`."""

    overlaps_acc = True    # ring_attention_backward may launch a step's kernels before asking for the step's sums

    def __init__(self, B, H, n, D, dtype, device, group=None, zigzag=False):
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.nb = 2 if zigzag else 1
        prev = (self.rank - 1) % self.world
        kv_shape = (2, 2, B, H, n, D)                        # [slot][K | V]
        acc_shape = (2, self.nb, 2, B, H, n // self.nb, D)   # [slot][chunk][dK | dV]
        self.kv_mem = symm_mem.empty(kv_shape, dtype=dtype, device=device)
        self.acc_mem = symm_mem.empty(acc_shape, dtype=torch.float32, device=device)
        self.kv_hdl = symm_mem.rendezvous(self.kv_mem, self.group)
        self.acc_hdl = symm_mem.rendezvous(self.acc_mem, self.group)
        self.kv_prev = self.kv_hdl.get_buffer(prev, kv_shape, dtype)
        self.acc_prev = self.acc_hdl.get_buffer(prev, acc_shape, torch.float32)
        self.side = torch.cuda.Stream(device)
        self.ev_kv = [torch.cuda.Event() for _ in range(2)]      # slot filled (by begin or by a pull)
        self.ev_kv_read = [torch.cuda.Event() for _ in range(2)]  # the step that computes on the slot has been enqueued
        self.ev_acc = [torch.cuda.Event() for _ in range(2)]     # accumulator slot pulled
        self.ev_acc_done = torch.cuda.Event()                    # this rank's additions of the last step are enqueued

    # channel 0: barriers on the compute stream; channel 1 / 2: the side stream's K / V and accumulator barriers
    def _barrier(self, channel=0):
        self.kv_hdl.barrier(channel=channel)

    def begin(self, K, V, chunk_shapes=None):
        self._barrier()                                       # nobody is still reading this rank's slots (previous call)
        self.kv_mem[0, 0].copy_(K), self.kv_mem[0, 1].copy_(V)
        self.ev_kv[0].record()
        self.want_acc = chunk_shapes is not None
        return [self.kv_mem[0, 0], self.kv_mem[0, 1]]

    def prefetch_kv(self, s):
        """Called before step s's kernels are enqueued: start pulling rank-1's shard of step s (= this rank's shard of
        step s+1) into the other slot."""
        if s + 1 >= self.world:
            return
        cur, nxt = s & 1, (s + 1) & 1
        compute = torch.cuda.current_stream()
        with torch.cuda.stream(self.side):
            if s == 0:
                self.side.wait_event(self.ev_kv[0])           # own slot 0 is filled (and every earlier kernel is done)
            else:
                self.side.wait_event(self.ev_kv_read[nxt])    # step s-1, which computed on slot `nxt`, is finished
            # every rank's slot `cur` is complete (its pull of step s-1 precedes this barrier on its side stream) and
            # nobody still pulls from this rank's slot `nxt` (rank+1 did so in its pull of step s-1)
            self._barrier(channel=1)
            self.kv_mem[nxt].copy_(self.kv_prev[cur])
            self.ev_kv[nxt].record(self.side)
        del compute

    def next_kv(self, s):
        cur, nxt = s & 1, (s + 1) & 1
        compute = torch.cuda.current_stream()
        self.ev_kv_read[cur].record(compute)                  # step s's kernels (enqueued above) read slot `cur`
        compute.wait_event(self.ev_kv[nxt])
        return [self.kv_mem[nxt, 0], self.kv_mem[nxt, 1]]

    def _acc_list(self, slot):
        return [self.acc_mem[slot, c, k] for c in range(self.nb) for k in range(2)]

    def first_acc(self):
        return self._acc_list(0)

    def prefetch_acc(self, s):
        """Start pulling the sums that arrive at step s (rank-1's slot of step s-1); called before step s's kernels."""
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.ev_acc_done)            # this rank's additions of step s-1 are complete ...
            self._barrier(channel=2)                          # ... and so are everybody else's
            self.acc_mem[s & 1].copy_(self.acc_prev[(s - 1) & 1])
            self.ev_acc[s & 1].record(self.side)

    def recv_acc(self, s):
        torch.cuda.current_stream().wait_event(self.ev_acc[s & 1])
        return self._acc_list(s & 1)

    def send_acc(self, s):
        # the sums stay in this rank's slot; rank+1 pulls them once this event (and its peers') has passed the barrier
        self.ev_acc_done.record(torch.cuda.current_stream())

    def final_acc(self):
        compute = torch.cuda.current_stream()
        home = torch.empty_like(self.acc_mem[0])
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.ev_acc_done)
            self._barrier(channel=2)
            home.copy_(self.acc_prev[(self.world - 1) & 1])
            done = torch.cuda.Event()
            done.record(self.side)
        compute.wait_event(done)
        home.record_stream(self.side)
        return [home[c, k] for c in range(self.nb) for k in range(2)]


def _check(Q, K, V, ops=None, zigzag=False, transport=None):
    """Everything that can be refused is refused HERE, before any communication is posted: an error raised mid-ring
    would leave the peers waiting on sends that never arrive."""
    if ops is CudaOps and Q.dtype not in (torch.float16, torch.bfloat16):
        raise TypeError(f"dtype {Q.dtype} not supported by the ring path (float16 / bfloat16 partials).")
    if ops is CudaOps and Q.dim() == 4 and Q.shape[-1] not in (64, 128):
        # fa_bwd_preprocess, fa_merge_partial and the attention kernels all run unpadded on the ring path
        raise ValueError(f"the ring path runs at head sizes 64 and 128 (got d={Q.shape[-1]}); pad the shards first")
    if transport is not None and getattr(transport, "nb", None) is not None and transport.nb != (2 if zigzag else 1):
        raise ValueError(f"transport was built with zigzag={transport.nb == 2} but the call says zigzag={bool(zigzag)}")
    if Q.dim() != 4 or Q.shape != K.shape or Q.shape != V.shape:
        raise ValueError("Q, K, V must all be local shards of shape (B, H, n, d) with the same n on every rank")
    if Q.dtype != K.dtype or K.dtype != V.dtype:
        raise ValueError("Q, K, V must have same dtype")


def _blocks(rank: int, world: int, n: int, zigzag: bool):
    """[(slice of the local sequence axis, global chunk id)].  Plain sharding: one chunk per rank, id = rank.  Zigzag:
    two half-length chunks per rank with ids rank and 2*world-1-rank, which balances the causal triangle (every rank
    does the same number of unmasked block products at every step)."""
    if not zigzag:
        return [(slice(0, n), rank)]
    if n % 2:
        raise ValueError("zigzag sharding needs an even local sequence length")
    return [(slice(0, n // 2), rank), (slice(n // 2, n), 2 * world - 1 - rank)]


def _relation(causal: bool, q_id: int, k_id: int):
    """None = fully masked (skip), True = the causal (diagonal) kernel, False = the unmasked kernel."""
    if not causal:
        return False
    return None if k_id > q_id else (k_id == q_id)


def ring_attention_forward(Q, K, V, causal: bool = False, softmax_scale: float = 1.0, group=None, ops=CudaOps,
                           zigzag: bool = False, transport=None):
    """Local shards in, (O_local in the input dtype, L_local (B,H,n,1) float32 in log2 units) out.
    zigzag=True: the local shard is [chunk rank | chunk 2G-1-rank] of a sequence cut into 2G chunks (causal balance).
    transport: None = torch.distributed point-to-point (NcclTransport(group)); a PeerTransport = NVLink peer-memory pulls."""
    _check(Q, K, V, ops, zigzag, transport)
    tr = transport if transport is not None else NcclTransport(group)
    r, G = tr.rank, tr.world
    B, H, n, d = Q.shape
    q_blocks = _blocks(r, G, n, zigzag)
    o_acc = [torch.empty((B, H, sl.stop - sl.start, d), dtype=torch.float32, device=Q.device) for sl, _ in q_blocks]
    l_acc = [torch.empty((B, H, sl.stop - sl.start), dtype=torch.float32, device=Q.device) for sl, _ in q_blocks]
    first = [True] * len(q_blocks)
    kv = tr.begin(K, V)
    for s in range(G):
        c = (r - s) % G                       # rank whose K / V shard is visiting at step s
        tr.prefetch_kv(s)                     # NCCL: next step's K / V are in flight while this step computes
        for qi, (qs, q_id) in enumerate(q_blocks):
            for ks, k_id in _blocks(c, G, n, zigzag):
                rel = _relation(causal, q_id, k_id)
                if rel is None:
                    continue
                o_part, l_part = ops.fwd(Q[:, :, qs], kv[0][:, :, ks], kv[1][:, :, ks], rel, float(softmax_scale))
                ops.merge(o_acc[qi], l_acc[qi], o_part, l_part.reshape(l_acc[qi].shape), first[qi])
                first[qi] = False
        if s + 1 < G:
            kv = tr.next_kv(s)
    O = torch.cat([ops.round(a, Q.dtype) for a in o_acc], dim=2) if len(o_acc) > 1 else ops.round(o_acc[0], Q.dtype)
    L = torch.cat(l_acc, dim=2) if len(l_acc) > 1 else l_acc[0]
    return O, L.unsqueeze(-1)


def ring_attention_backward(Q, K, V, O, dO, L, causal: bool = False, softmax_scale: float = 1.0, group=None, ops=CudaOps,
                            zigzag: bool = False, transport=None):
    """Local shards (and the forward's O_local, L_local) in, (dQ, dK, dV) of the local shards out."""
    _check(Q, K, V, ops, zigzag, transport)
    tr = transport if transport is not None else NcclTransport(group)
    r, G = tr.rank, tr.world
    B, H, n, d = Q.shape
    q_blocks = _blocks(r, G, n, zigzag)
    nb = len(q_blocks)
    L3 = L.reshape(B, H, n).float()
    delta = ops.delta(O, dO)
    Lb = [L3[:, :, sl].contiguous() for sl, _ in q_blocks]
    db = [delta[:, :, sl].contiguous() for sl, _ in q_blocks]
    shapes = [(B, H, sl.stop - sl.start, d) for sl, _ in q_blocks]
    dq_acc = [torch.empty(shp, dtype=torch.float32, device=Q.device) for shp in shapes]
    first_q = [True] * nb
    kv = tr.begin(K, V, shapes)
    dkv_acc = tr.first_acc()                  # travelling accumulators: [chunk][dK | dV] of the visiting shard
    overlap = bool(getattr(tr, "overlaps_acc", False))
    for s in range(G):
        c = (r - s) % G
        tr.prefetch_kv(s)
        if s > 0:
            if overlap:
                tr.prefetch_acc(s)            # the arriving sums travel while this step's kernels run
            else:
                dkv_acc = tr.recv_acc(s)      # sums of shard c so far, arriving with it from rank r-1
        parts = []
        for qi, (qs, q_id) in enumerate(q_blocks):
            for ki, (ks, k_id) in enumerate(_blocks(c, G, n, zigzag)):
                rel = _relation(causal, q_id, k_id)
                if rel is None:
                    continue
                parts.append((qi, ki) + tuple(ops.bwd(Q[:, :, qs], kv[0][:, :, ks], kv[1][:, :, ks], O[:, :, qs],
                                                      dO[:, :, qs], Lb[qi], rel, float(softmax_scale), db[qi])))
        if s > 0 and overlap:
            dkv_acc = tr.recv_acc(s)
        touched = [s > 0] * nb                # at s = 0 the accumulators of the own shard start from nothing
        for qi, ki, dq_p, dk_p, dv_p in parts:   # same order of additions as the kernels were launched in: deterministic
            ops.accumulate(dq_acc[qi], dq_p, first_q[qi])
            first_q[qi] = False
            ops.accumulate(dkv_acc[2 * ki], dk_p, not touched[ki])
            ops.accumulate(dkv_acc[2 * ki + 1], dv_p, not touched[ki])
            touched[ki] = True
        for ki in range(nb):                  # a chunk nobody on this rank attends to at s = 0 still needs defined sums
            if not touched[ki]:
                dkv_acc[2 * ki].zero_(), dkv_acc[2 * ki + 1].zero_()
        tr.send_acc(s)                        # on to rank r+1 (the last hop brings every shard's sums home)
        if s + 1 < G:
            kv = tr.next_kv(s)
    home = tr.final_acc()
    cat = lambda parts: torch.cat(parts, dim=2) if len(parts) > 1 else parts[0]
    dQ = cat([ops.round(a, Q.dtype) for a in dq_acc])
    dK = cat([ops.round(home[2 * ki], Q.dtype) for ki in range(nb)])
    dV = cat([ops.round(home[2 * ki + 1], Q.dtype) for ki in range(nb)])
    return dQ, dK, dV


class RingAttention(torch.autograd.Function):
    """O_local = RingAttention.apply(Q_local, K_local, V_local, causal, softmax_scale, group, zigzag, transport)."""

    @staticmethod
    def forward(ctx, Q, K, V, causal=False, softmax_scale=1.0, group=None, zigzag=False, transport=None):
        O, L = ring_attention_forward(Q, K, V, causal, softmax_scale, group, zigzag=zigzag, transport=transport)
        ctx.save_for_backward(Q, K, V, O, L)
        ctx.causal, ctx.softmax_scale, ctx.group, ctx.zigzag = bool(causal), float(softmax_scale), group, bool(zigzag)
        ctx.transport = transport
        return O

    @staticmethod
    def backward(ctx, dO):
        Q, K, V, O, L = ctx.saved_tensors
        dQ, dK, dV = ring_attention_backward(Q, K, V, O, dO, L, ctx.causal, ctx.softmax_scale, ctx.group,
                                             zigzag=ctx.zigzag, transport=ctx.transport)
        return dQ, dK, dV, None, None, None, None, None


def zigzag_shard(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """The local shard of a full (B, H, N, D) tensor under zigzag sharding: chunks rank and 2*world-1-rank of 2*world."""
    c = t.shape[2] // (2 * world)
    return torch.cat([t[:, :, rank * c:(rank + 1) * c], t[:, :, (2 * world - 1 - rank) * c:(2 * world - rank) * c]], dim=2)
