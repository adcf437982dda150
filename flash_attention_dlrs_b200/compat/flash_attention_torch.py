"""Flat-module shim: put this directory on sys.path and the reference's scripts
(`from flash_attention_torch import FlashAttention, FlashAttentionDeterministic`, src/test_torch.py:2)
resolve to the B200 implementation unchanged."""
from flash_attention_dlrs_b200.flash_attention_torch import *  # noqa: F401,F403
from flash_attention_dlrs_b200.flash_attention_torch import (  # noqa: F401
    MIN_TENSOR_SIZE, FlashAttention, FlashAttentionDeterministic, convert_triton_dtype, flash_attention)
