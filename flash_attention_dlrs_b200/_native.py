"""Torch-tensor front of the C ABI: allocation, stride hygiene, head-size padding, launches on torch's current stream.

Mirrors what the reference's torch boundary does around its Triton launches (flash_attention_torch.py:38-74,
96-154): pad the head size, allocate outputs with torch so the caching allocator owns every buffer, pass all
four strides of every tensor.  Everything here runs on the GPU through libfa_b200.so; there is no fallback.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

_DTYPES = {torch.float16: _lib.FA_DTYPE_F16, torch.bfloat16: _lib.FA_DTYPE_BF16, torch.float32: _lib.FA_DTYPE_F32,
           # forward only; float8_e5m2 is the FP8 type of the reference's dtype map (flash_attention_torch.py:15-16)
           torch.float8_e5m2: _lib.FA_DTYPE_F8E5M2, torch.float8_e4m3fn: _lib.FA_DTYPE_F8E4M3}
FP8_DTYPES = (torch.float8_e5m2, torch.float8_e4m3fn)
MAX_HEAD_DIM = 128


def dtype_code(dtype: torch.dtype) -> int:
    try:
        return _DTYPES[dtype]
    except KeyError:
        raise TypeError(f"dtype {dtype} not supported.") from None


def padded_head_dim(d: int, dtype: torch.dtype) -> int:
    """Head size the kernels run at: 64 / 128 for 16-bit inputs, a power of two in [16, 128] for float32
    (the reference pads to max(next_pow2(d), 16), flash_attention_torch.py:38)."""
    if d < 1 or d > MAX_HEAD_DIM:
        raise ValueError(f"head size d={d} not supported (1 <= d <= {MAX_HEAD_DIM})")
    if dtype == torch.float32:
        return max(1 << (d - 1).bit_length(), 16)
    if dtype in FP8_DTYPES:
        return 128   # one 128-byte TMA box per row
    return 64 if d <= 64 else 128


def _kernel_ready(t: torch.Tensor) -> torch.Tensor:
    """A view the kernels can address: unit inner stride, 16-byte aligned base and outer strides."""
    if t.is_contiguous() and t.data_ptr() % 16 == 0 and (t.shape[-1] * t.element_size()) % 16 == 0:
        return t   # the common case, without the stride arithmetic (this function is on the launch path)
    gran = 16 // t.element_size()
    B, H, N, _ = t.shape
    sB, sH, sN, sD = t.stride()
    ok = sD == 1 and t.data_ptr() % 16 == 0 and sN % gran == 0
    ok = ok and (H == 1 or sH % gran == 0) and (B == 1 or sB % gran == 0)
    # broadcast (stride 0) or reversed views cannot be described to TMA: copy them
    ok = ok and sN > 0 and (H == 1 or sH > 0) and (B == 1 or sB > 0) and (N == 1 or sN >= t.shape[-1])
    return t if ok else t.contiguous()


def _pad_d(t: torch.Tensor, d_run: int) -> torch.Tensor:
    d = t.shape[-1]
    if d == d_run:
        return t
    if t.dtype in FP8_DTYPES:   # F.pad has no FP8 kernel; the all-zero byte is +0.0 in both formats
        out = torch.zeros((*t.shape[:-1], d_run), dtype=torch.uint8, device=t.device)
        out[..., :d] = t.view(torch.uint8)
        return out.view(t.dtype)
    return torch.nn.functional.pad(t, (0, d_run - d), mode="constant", value=0.0)


def _stream_ptr(dev: torch.device) -> ctypes.c_void_p:
    """torch's current stream on `dev` as a cudaStream_t (the raw getter: this is on the launch path of every call)."""
    idx = dev.index if dev.index is not None else torch._C._cuda_getDevice()
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(idx))


class _on_device:
    """`with torch.cuda.device(dev)` that costs nothing when `dev` is already current (the usual case)."""

    __slots__ = ("guard",)

    def __init__(self, dev: torch.device):
        idx = dev.index
        self.guard = None if idx is None or idx == torch._C._cuda_getDevice() else torch.cuda.device(idx)

    def __enter__(self):
        if self.guard is not None:
            self.guard.__enter__()

    def __exit__(self, *exc):
        if self.guard is not None:
            self.guard.__exit__(*exc)


def _ptr(t: torch.Tensor) -> ctypes.c_void_p:
    return ctypes.c_void_p(t.data_ptr())


def _seqlens_arg(seqlens, B: int, device):
    """(tensor kept alive, c_void_p) for the optional per-batch valid lengths: int32, contiguous, on `device`."""
    if seqlens is None:
        return None, ctypes.c_void_p(0)
    t = torch.as_tensor(seqlens, device=device).to(torch.int32).contiguous()
    if t.dim() != 1 or t.shape[0] != B:
        raise ValueError(f"seqlens must have shape (B,) = ({B},), got {tuple(t.shape)}")
    return t, _ptr(t)


def dropout_threshold(dropout_p: float) -> int:
    """The kernels' quantisation of a dropout probability: an entry is dropped with probability threshold / 256
    (fa_dropout.cuh); 0 = no dropout."""
    p = float(dropout_p)
    if not 0.0 <= p < 1.0:
        raise ValueError(f"dropout_p must be in [0, 1), got {dropout_p}")
    return min(int(p * 256.0 + 0.5), 255)


def _dropout_args(dropout_p, dropout_seed):
    p = float(dropout_p)
    if dropout_threshold(p) == 0:
        return 0.0, 0
    if dropout_seed is None:
        raise ValueError("dropout_p > 0 needs a dropout_seed (the backward regenerates the mask from it)")
    return p, int(dropout_seed) & 0xFFFFFFFFFFFFFFFF


class AttentionMask:
    """An arbitrary attention mask in the layout the kernels read (include/fa_b200.h, `fa_attn_mask`): one BIT per
    (query, key), 1 = attend, 16 bytes per row and 128-key block — once as [.., query, key / 8] (forward, dQ kernel)
    and once transposed [.., key, query / 8] (dK/dV kernel) — plus a byte per 128 x 128 block saying whether the block
    is empty, mixed or fully visible.  Build it once and pass it as `attn_mask` to reuse the packed copies across calls;
    a plain bool tensor is wrapped on the fly.

    `mask`: bool (or any dtype, non-zero = attend) of shape (N, N), (B|1, N, N) or (B|1, H|1, N, N)."""

    def __init__(self, mask: torch.Tensor):
        self.window = None
        if mask.dim() == 2:
            mask = mask[None, None]
        elif mask.dim() == 3:
            mask = mask[:, None]
        if mask.dim() != 4 or mask.shape[-1] != mask.shape[-2]:
            raise ValueError(f"attn_mask must be (N, N), (B, N, N) or (B, H, N, N), got {tuple(mask.shape)}")
        m = mask if mask.dtype == torch.bool else (mask != 0)
        Bm, Hm, N, _ = m.shape
        pitch = (N + 127) // 128 * 128

        weights = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.int16, device=m.device)

        def pack(x):   # (.., N, N) bool -> (.., N, pitch / 8) uint8, entry j = bit (j & 7) of byte (j >> 3)
            buf = torch.zeros((Bm, Hm, N, pitch), dtype=torch.int16, device=x.device)
            buf[..., :N] = x
            return (buf.view(Bm, Hm, N, pitch // 8, 8) * weights).sum(-1).to(torch.uint8).contiguous()

        self.rows, self.cols = pack(m), pack(m.transpose(-1, -2))
        # 128 x 128 block summary: 0 = nothing of block (i, j) is visible (the kernels skip it), 2 = everything is
        # (its mask bytes are not read), 1 = mixed
        nb = pitch // 128
        sq = torch.zeros((Bm, Hm, pitch, pitch), dtype=torch.bool, device=m.device)
        sq[..., :N, :N] = m
        sq = sq.view(Bm, Hm, nb, 128, nb, 128)
        self.blocks = (sq.any(-1).any(-2).to(torch.uint8) + sq.all(-1).all(-2).to(torch.uint8)).contiguous()
        self.shape = (Bm, Hm, N)
        self._struct = None

    @classmethod
    def sliding_window(cls, N: int, left: int, right: int, device="cuda"):
        """Mask of a local-attention band: query i sees keys i - left .. i + right (combine with causal=True, or
        right = 0, for a causal window).  With block skipping the kernels' time is proportional to the band.
        No mask bytes are stored: the kernels compute the band from the two numbers, and the 128 x 128 block summary
        (which blocks to skip, which are fully visible) is computed here from them — O((N / 128)^2) memory."""
        left, right = int(left), int(right)
        if left < 0 or right < 0:
            raise ValueError("sliding_window needs left, right >= 0")
        self = cls.__new__(cls)
        self.window = (left, right)
        self.rows = self.cols = None
        nb = (N + 127) // 128
        bi = torch.arange(nb, device=device)
        # key - query over block (i, j) ranges over [128 (j - i) - 127, 128 (j - i) + 127]
        dmin = 128 * (bi[None, :] - bi[:, None]) - 127
        dmax = dmin + 254
        some = (dmax >= -left) & (dmin <= right)
        full = (dmin >= -left) & (dmax <= right)
        self.blocks = (some.to(torch.uint8) + full.to(torch.uint8))[None, None].contiguous()
        self.shape = (1, 1, N)
        self._struct = None
        return self

    def struct(self, B: int, H: int, N: int, device):
        """ctypes fa_attn_mask for a (B, H, N, .) problem; size-1 batch / head dims broadcast (stride 0)."""
        Bm, Hm, Nm = self.shape
        if Nm != N or Bm not in (1, B) or Hm not in (1, H):
            raise ValueError(f"attn_mask of shape {(Bm, Hm, Nm, Nm)} does not broadcast to {(B, H, N, N)}")
        if self.blocks.device != device:
            raise ValueError(f"attn_mask is on {self.blocks.device}, the inputs on {device}")
        if self._struct is None:
            st = _lib.AttnMaskStruct()
            st.window_left, st.window_right = self.window if self.window is not None else (-1, -1)
            for name, t in (("rows", self.rows), ("cols", self.cols), ("blocks", self.blocks)):
                if t is None:
                    continue
                sB, sH, sN, _ = t.stride()
                setattr(st, name, t.data_ptr())
                setattr(st, name + "_strides", _lib._I64x3(0 if Bm == 1 else sB, 0 if Hm == 1 else sH, sN))
            self._struct = st
        return self._struct


def _mask_arg(attn_mask, B, H, N, device):
    """(AttentionMask kept alive, pointer to its fa_attn_mask or NULL)"""
    if attn_mask is None:
        return None, None
    am = attn_mask if isinstance(attn_mask, AttentionMask) else AttentionMask(attn_mask)
    return am, ctypes.byref(am.struct(B, H, N, device))


def forward(Q: torch.Tensor, K: torch.Tensor, V: torch.Tensor, causal: bool, softmax_scale: float, out=None,
            peer_ptrs=(), seqlens=None, dropout_p: float = 0.0, dropout_seed=None, attn_mask=None):
    """O (B,H,N,d) in the input dtype and L (B,H,N) float32 in log2 units.  Inputs already validated.

    `out` = (data_ptr, element strides (sB, sH, sN, 1)) makes the kernel write O at a caller-owned address (a window of
    a gathered buffer, possibly an NVLS multicast address) instead of a fresh tensor, and `peer_ptrs` (<= 7 device
    pointers) adds peer-mapped copies with the same strides (fa_fwd_peers); both need an unpadded head size and return
    O = None.

    `seqlens` (B,) int: key-padding mask — batch element b has seqlens[b] valid tokens; keys beyond are masked out and
    the rows of O / L beyond are zero.

    `dropout_p`, `dropout_seed`: in-kernel dropout of the attention probabilities (16-bit and float32; quantised to
    dropout_threshold(p) / 256, kept entries scaled by the reciprocal of the keep rate).  The mask is a pure function
    of (seed, b, h, query, key) — oracle/attention_oracle.py restates it; L is that of the undropped scores.

    `attn_mask`: arbitrary mask, bool (N, N) / (B|1, N, N) / (B|1, H|1, N, N) or an AttentionMask, True = attend, ANDed
    with `causal` and `seqlens`; a query with no visible key gets O = 0, L = -inf.  16-bit and float32."""
    lib = _lib.load()
    B, H, N, d = Q.shape
    code = dtype_code(Q.dtype)
    drop_p, drop_seed = _dropout_args(dropout_p, dropout_seed)
    if drop_p and Q.dtype in FP8_DTYPES:
        raise TypeError(f"dropout is not implemented for the FP8 forward ({Q.dtype})")
    if attn_mask is not None and Q.dtype in FP8_DTYPES:
        raise TypeError(f"attention masks are not implemented for the FP8 forward ({Q.dtype})")
    am, am_args = _mask_arg(attn_mask, B, H, N, Q.device)
    if Q.numel() == 0 and out is None:   # empty batch / no heads / no tokens: nothing to launch
        return torch.empty_like(Q), torch.empty((B, H, N), dtype=torch.float32, device=Q.device)
    d_run = padded_head_dim(d, Q.dtype)
    q, k, v = (_kernel_ready(_pad_d(t, d_run)) for t in (Q, K, V))
    sl, sl_ptr = _seqlens_arg(seqlens, B, Q.device)
    new = torch.empty if sl is None else torch.zeros   # padded rows are not written by the kernels
    L = new((B, H, N), dtype=torch.float32, device=Q.device)
    if out is None:
        O = (new((B, H, N, d_run), dtype=Q.dtype, device=Q.device) if Q.dtype not in FP8_DTYPES or sl is None
             else torch.zeros((B, H, N, d_run), dtype=torch.uint8, device=Q.device).view(Q.dtype))
        o_ptr, o_strides = _ptr(O), _lib.strides4(O)
    else:
        if d_run != d:
            raise ValueError(f"writing O in place needs a head size the kernels run at (d={d} is padded to {d_run})")
        O = None
        o_ptr, o_strides = ctypes.c_void_p(int(out[0])), _lib._I64x4(*out[1])
    peers = (ctypes.c_void_p * max(len(peer_ptrs), 1))(*[int(x) for x in peer_ptrs])
    with _on_device(Q.device):
        rc = lib.fa_fwd_peers(_ptr(q), _ptr(k), _ptr(v), o_ptr, _ptr(L), B, H, N, d_run,
                              _lib.strides4(q), _lib.strides4(k), _lib.strides4(v), o_strides,
                              code, float(softmax_scale), int(bool(causal)), len(peer_ptrs), peers, sl_ptr,
                              drop_p, drop_seed, am_args, _stream_ptr(Q.device))
    _lib.check(rc, "fa_fwd")
    if O is None:
        return None, L
    return (O if d_run == d else O[..., :d]), L


def backward_preprocess(O: torch.Tensor, dO: torch.Tensor) -> torch.Tensor:
    """delta (B,H,N) float32 = rowsum(O * dO).  O and dO must have the same (kernel-legal) head size."""
    lib = _lib.load()
    B, H, N, d = O.shape
    code = dtype_code(O.dtype)
    o, do = _kernel_ready(O), _kernel_ready(dO)
    delta = torch.empty((B, H, N), dtype=torch.float32, device=O.device)
    with _on_device(O.device):
        rc = lib.fa_bwd_preprocess(_ptr(o), _ptr(do), _ptr(delta), B, H, N, d,
                                   _lib.strides4(o), _lib.strides4(do), code, _stream_ptr(O.device))
    _lib.check(rc, "fa_bwd_preprocess")
    return delta


BWD_DKDV, BWD_DQ, BWD_FUSED = 1, 2, 4


def backward(Q, K, V, O, dO, L, causal: bool, softmax_scale: float, which: int | None = None, delta=None,
             seqlens=None, dropout_p: float = 0.0, dropout_seed=None, attn_mask=None):
    """dQ, dK, dV (B,H,N,d) in the input dtype; deterministic (bit-identical across runs).
    `which` = None runs what fa_bwd runs: the two-kernel path (BWD_DKDV | BWD_DQ; either half can be selected alone,
    unselected outputs are uninitialised).  BWD_FUSED selects the single-pass kernel with the ordered dQ reduction
    (16-bit inputs only).  `delta` may carry a precomputed rowsum(O * dO) to skip the preprocess launch.
    `seqlens` as in forward(): gradient rows beyond seqlens[b] are zero (two-kernel path only).
    `dropout_p`, `dropout_seed`, `attn_mask`: the values the forward ran with (two-kernel path only)."""
    lib = _lib.load()
    B, H, N, d = Q.shape
    code = dtype_code(Q.dtype)
    drop_p, drop_seed = _dropout_args(dropout_p, dropout_seed)
    am, am_args = _mask_arg(attn_mask, B, H, N, Q.device)
    if Q.dtype in FP8_DTYPES:
        raise TypeError(f"dtype {Q.dtype} not supported in backward (the FP8 path is forward-only).")
    if Q.numel() == 0:
        return torch.empty_like(Q), torch.empty_like(K), torch.empty_like(V)
    d_run = padded_head_dim(d, Q.dtype)
    q, k, v, o, do = (_kernel_ready(_pad_d(t, d_run)) for t in (Q, K, V, O, dO))
    lse = L if (L.dtype == torch.float32 and L.is_contiguous()) else L.to(torch.float32).contiguous()  # (B,H,N[,1])
    if delta is None:
        delta = backward_preprocess(o, do)
    sl, sl_ptr = _seqlens_arg(seqlens, B, Q.device)
    new = torch.empty if sl is None else torch.zeros   # padded rows are not written by the kernels
    # three independent allocations: a caller that keeps only one gradient alive must not pin the other two
    dQ, dK, dV = (new((B, H, N, d_run), dtype=Q.dtype, device=Q.device) for _ in range(3))
    if which is None:
        which = BWD_DKDV | BWD_DQ
    ws_bytes = lib.fa_bwd_workspace_bytes(B, H, N, d_run, code, int(bool(causal)), int(which))
    # scratch of the ordered dQ reduction, owned by the caching allocator like every other buffer (the reference
    # allocates its dQ lock buffers the same way, flash_attention_torch.py:107-109)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=Q.device) if ws_bytes else None
    with _on_device(Q.device):
        rc = lib.fa_bwd_partial(_ptr(q), _ptr(k), _ptr(v), _ptr(do), _ptr(lse), _ptr(delta), _ptr(dQ), _ptr(dK),
                                _ptr(dV), _ptr(ws) if ws is not None else ctypes.c_void_p(0), ws_bytes, B, H, N,
                                d_run, _lib.strides4(q), _lib.strides4(k), _lib.strides4(v), _lib.strides4(do),
                                _lib.strides4(dQ), _lib.strides4(dK), _lib.strides4(dV),
                                code, float(softmax_scale), int(bool(causal)), int(which), sl_ptr,
                                drop_p, drop_seed, am_args, _stream_ptr(Q.device))
    _lib.check(rc, "fa_bwd")
    if d_run != d:
        dQ, dK, dV = dQ[..., :d], dK[..., :d], dV[..., :d]
    return dQ, dK, dV


def forward_rect(Q: torch.Tensor, K: torch.Tensor, V: torch.Tensor, softmax_scale: float):
    """Rectangular attention: Q (B,H,Nq,d) against K, V (B,H,Nk,d), Nk != Nq allowed (fa_fwd_rect: float16 / bfloat16,
    d in {64, 128}, non-causal).  Returns O (B,H,Nq,d) and L (B,H,Nq) float32 in log2 units."""
    lib = _lib.load()
    B, H, Nq, d = Q.shape
    Nk = K.shape[2]
    if K.shape != V.shape or K.shape[:2] != Q.shape[:2] or K.shape[3] != d:
        raise ValueError("forward_rect: Q (B,H,Nq,d), K and V (B,H,Nk,d)")
    q, k, v = (_kernel_ready(t) for t in (Q, K, V))
    O = torch.empty((B, H, Nq, d), dtype=Q.dtype, device=Q.device)
    L = torch.empty((B, H, Nq), dtype=torch.float32, device=Q.device)
    with _on_device(Q.device):
        rc = lib.fa_fwd_rect(_ptr(q), _ptr(k), _ptr(v), _ptr(O), _ptr(L), B, H, Nq, Nk, d, _lib.strides4(q),
                             _lib.strides4(k), _lib.strides4(v), _lib.strides4(O), dtype_code(Q.dtype),
                             float(softmax_scale), _stream_ptr(Q.device))
    _lib.check(rc, "fa_fwd_rect")
    return O, L


def backward_rect(Q, K, V, O, dO, L, softmax_scale: float, delta=None):
    """Gradients of forward_rect: dQ (B,H,Nq,d), dK, dV (B,H,Nk,d); deterministic (the two-kernel path)."""
    lib = _lib.load()
    B, H, Nq, d = Q.shape
    Nk = K.shape[2]
    q, k, v, o, do = (_kernel_ready(t) for t in (Q, K, V, O, dO))
    lse = L if (L.dtype == torch.float32 and L.is_contiguous()) else L.to(torch.float32).contiguous()
    if delta is None:
        delta = backward_preprocess(o, do)
    dQ = torch.empty((B, H, Nq, d), dtype=Q.dtype, device=Q.device)
    dK, dV = (torch.empty((B, H, Nk, d), dtype=Q.dtype, device=Q.device) for _ in range(2))
    with _on_device(Q.device):
        rc = lib.fa_bwd_rect(_ptr(q), _ptr(k), _ptr(v), _ptr(do), _ptr(lse), _ptr(delta), _ptr(dQ), _ptr(dK), _ptr(dV),
                             B, H, Nq, Nk, d, _lib.strides4(q), _lib.strides4(k), _lib.strides4(v), _lib.strides4(do),
                             _lib.strides4(dQ), _lib.strides4(dK), _lib.strides4(dV), dtype_code(Q.dtype),
                             float(softmax_scale), _stream_ptr(Q.device))
    _lib.check(rc, "fa_bwd_rect")
    return dQ, dK, dV
