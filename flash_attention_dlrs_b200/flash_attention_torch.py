"""torch.autograd.Function boundary of the attention path — same names and argument meaning as the reference's
flash_attention_torch.py (FlashAttention :21, FlashAttentionDeterministic :161, convert_triton_dtype :7).

    O = FlashAttention.apply(Q, K, V)                       # reference call, scale = 1.0, non-causal
    O = FlashAttention.apply(Q, K, V, causal, softmax_scale)  # added trailing arguments, order as in the
                                                              # vendored tutorial (_attention.forward :441)
    O = FlashAttention.apply(Q, K, V, causal, softmax_scale, seqlens)   # + per-batch valid lengths (key padding)
    O = FlashAttention.apply(Q, K, V, causal, softmax_scale, seqlens, dropout_p[, dropout_seed])   # + in-kernel dropout
    O = FlashAttention.apply(Q, K, V, causal, softmax_scale, None, 0.0, None, attn_mask)             # + arbitrary bool mask

Differences from the reference, all deliberate:
  * the kernels are hand-written sm_100a CUDA reached through libfa_b200.so (no Triton, no autotune);
  * the backward is deterministic, so FlashAttentionDeterministic is the same Function;
  * L (base-2 logsumexp, flash_attention_kernels.py:106) is kept in float32 instead of the input dtype;
  * bfloat16 is accepted; float8_e5m2 (in the reference's dtype map) and float8_e4m3fn run the forward only (P is cast
    to the FP8 type before P.V, as the tutorial's fp8 path does, flash_attention_openai_tutorial.py:66-67); float64 is not;
  * head sizes that need padding also work in backward (the reference's padded backward is broken).
"""
from __future__ import annotations

import os

import torch

from . import _native

MIN_TENSOR_SIZE = 16  # flash_attention_torch.py:5

# The plain call FlashAttention.apply(Q, K, V[, causal, softmax_scale]) — the only one the reference has — is served by a
# C++ autograd node over the same C-ABI entry points when the in-tree extension is built (csrc/torch_binding.cpp): no
# interpreter, ctypes marshalling or GIL hand-over on the launch path, which is what bounds a step below N ~ 1024.
# Everything else (seqlens, dropout, masks, padded head sizes, FP8, errors) runs the Python Function below.  Both call the
# same kernels with the same arguments: results are bit-identical (tests/test_gpu_parity.py).  Not used with FA_B200_LIB
# (an alternative build of the library) or FA_B200_NO_CPP_NODE=1.
USE_CPP_NODE = True
_cpp_node_cache = []


def _cpp_node():
    if not _cpp_node_cache:
        node = None
        if not os.environ.get("FA_B200_LIB") and os.environ.get("FA_B200_NO_CPP_NODE") != "1":
            try:
                from . import _lib
                _lib.load()   # the library first: a missing libfa_b200.so must fail with the loader's message
                from . import _fa_torch as node
            except ImportError:
                node = None
        _cpp_node_cache.append(node)
    return _cpp_node_cache[0]


def convert_triton_dtype(torch_dtype):
    """dtype whitelist with the reference's error (flash_attention_torch.py:7-18); returns the C-ABI dtype code."""
    return _native.dtype_code(torch_dtype)


def _validate(Q, K, V):
    dev = Q.device
    if dev.type != "cuda" or dev != K.device or dev != V.device:
        raise NotImplementedError("Q, K, V must be on the same CUDA device")
    if Q.dim() != 4 or Q.shape != K.shape or Q.shape != V.shape:
        raise ValueError("Q, K, V must all be of shape (B, H, N, d)")
    if Q.dtype != K.dtype or K.dtype != V.dtype:
        raise ValueError("Q, K, V must have same dtype")
    convert_triton_dtype(Q.dtype)


class FlashAttention(torch.autograd.Function):
    @classmethod
    def apply(cls, Q, K, V, causal=False, softmax_scale=1.0, seqlens=None, dropout_p=0.0, dropout_seed=None,
              attn_mask=None):
        if USE_CPP_NODE and seqlens is None and attn_mask is None and not dropout_p:
            node = _cpp_node()
            if (node is not None and isinstance(Q, torch.Tensor) and isinstance(K, torch.Tensor)
                    and isinstance(V, torch.Tensor) and node.supported(Q, K, V)):
                return node.flash_attention(Q, K, V, bool(causal), float(softmax_scale))
        return super().apply(Q, K, V, causal, softmax_scale, seqlens, dropout_p, dropout_seed, attn_mask)

    @staticmethod
    def forward(ctx, Q: torch.Tensor, K: torch.Tensor, V: torch.Tensor, causal: bool = False,
                softmax_scale: float = 1.0, seqlens=None, dropout_p: float = 0.0, dropout_seed=None,
                attn_mask=None) -> torch.Tensor:
        _validate(Q, K, V)
        causal = bool(causal)
        softmax_scale = float(softmax_scale)
        dropout_p = float(dropout_p)
        if _native.dropout_threshold(dropout_p) and dropout_seed is None:
            # a fresh mask per call, reproducible under torch.manual_seed (drawn from torch's default CPU generator)
            dropout_seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        if attn_mask is not None and not isinstance(attn_mask, _native.AttentionMask):
            attn_mask = _native.AttentionMask(attn_mask)   # packed once, reused by the backward
        O, L = _native.forward(Q, K, V, causal, softmax_scale, seqlens=seqlens, dropout_p=dropout_p,
                               dropout_seed=dropout_seed, attn_mask=attn_mask)   # runs without graph recording
        # same saved set as the reference (flash_attention_torch.py:77), unpadded
        ctx.save_for_backward(Q, K, V, O, L)
        ctx.causal = causal
        ctx.softmax_scale = softmax_scale
        ctx.seqlens = seqlens
        ctx.dropout = (dropout_p, dropout_seed)
        ctx.attn_mask = attn_mask
        return O

    @staticmethod
    def backward(ctx, grad_outputs, *args):
        Q, K, V, O, L = ctx.saved_tensors
        dO = grad_outputs
        if Q.dtype != dO.dtype:
            raise ValueError("dO must have same dtype as inputs")
        dQ, dK, dV = _native.backward(Q, K, V, O, dO, L, ctx.causal, ctx.softmax_scale, seqlens=ctx.seqlens,
                                      dropout_p=ctx.dropout[0], dropout_seed=ctx.dropout[1], attn_mask=ctx.attn_mask)
        return dQ, dK, dV, None, None, None, None, None, None


# The reference's second Function differs only in which (broken) backward kernel it launches
# (flash_attention_torch.py:247,274); the backward here is always deterministic.
FlashAttentionDeterministic = FlashAttention


def flash_attention(Q, K, V, causal: bool = False, softmax_scale: float = 1.0, seqlens=None, dropout_p: float = 0.0,
                    dropout_seed=None, attn_mask=None) -> torch.Tensor:
    """Keyword-friendly front of FlashAttention.apply.  `seqlens` (B,) int: per-batch valid length (key-padding mask,
    the "masking" of the reference's roadmap, README.md:35-37); rows beyond it are zero in O and in the gradients.
    `dropout_p`: in-kernel dropout of the attention probabilities (the "dropout" of the same roadmap), quantised to
    1/256; `dropout_seed` fixes the mask (default: drawn from torch's CPU generator).
    `attn_mask`: arbitrary mask, bool (N, N) / (B|1, N, N) / (B|1, H|1, N, N) or a prepared `AttentionMask`, True = attend,
    ANDed with `causal` and `seqlens`; queries with no visible key get O = 0 and zero gradients."""
    return FlashAttention.apply(Q, K, V, causal, softmax_scale, seqlens, dropout_p, dropout_seed, attn_mask)
