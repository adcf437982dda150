"""Functional (non-autograd) forward / backward — same names and argument order as the reference's
flash_attention_wrappers.py (flash_attention_forward :7, flash_attention_backward :66).

    O, L = flash_attention_forward(Q, K, V, dev[, causal, softmax_scale])
    dQ, dK, dV = flash_attention_backward(Q, K, V, O, dO, L, dev[, deterministic, causal, softmax_scale])

L is (B, H, N, 1) like the reference's (flash_attention_wrappers.py:38) but float32, in log2 units:
L = log2(e) * logsumexp_j(softmax_scale * S_ij)  (flash_attention_kernels.py:106).  `deterministic` is accepted
for signature compatibility and ignored: the backward is always deterministic.  `seqlens` (B,) int, optional, last:
per-batch valid length (key-padding mask); rows beyond it are zero in O, L and the gradients.  `dropout_p`,
`dropout_seed` (after it): in-kernel dropout; `attn_mask` (after those): arbitrary bool mask or a prepared AttentionMask;
give the backward the values the forward ran with.
"""
from __future__ import annotations

import torch

from . import _native
from .flash_attention_torch import convert_triton_dtype


def _check(Q, K, V, dev):
    # the reference uses bare asserts here (flash_attention_wrappers.py:20-22, 77-79)
    assert Q.dim() == 4
    assert Q.shape == K.shape and K.shape == V.shape
    assert Q.dtype == K.dtype and K.dtype == V.dtype
    convert_triton_dtype(Q.dtype)
    dev = torch.device(dev)
    if dev.type != "cuda" or any(t.device.type != "cuda" for t in (Q, K, V)):
        raise NotImplementedError("Q, K, V must be on the same CUDA device")


def flash_attention_forward(Q, K, V, dev, causal: bool = False, softmax_scale: float = 1.0, seqlens=None,
                            dropout_p: float = 0.0, dropout_seed=None, attn_mask=None):
    _check(Q, K, V, dev)
    O, L = _native.forward(Q, K, V, bool(causal), float(softmax_scale), seqlens=seqlens, dropout_p=dropout_p,
                           dropout_seed=dropout_seed, attn_mask=attn_mask)
    return O, L.unsqueeze(-1)


def flash_attention_backward(Q, K, V, O, dO, L, dev, deterministic: bool = False, causal: bool = False,
                             softmax_scale: float = 1.0, seqlens=None, dropout_p: float = 0.0, dropout_seed=None,
                             attn_mask=None):
    _check(Q, K, V, dev)
    assert O.shape == Q.shape and dO.shape == Q.shape
    assert dO.dtype == Q.dtype and O.dtype == Q.dtype
    return _native.backward(Q, K, V, O, dO, L, bool(causal), float(softmax_scale), seqlens=seqlens,
                            dropout_p=dropout_p, dropout_seed=dropout_seed, attn_mask=attn_mask)
