"""B200-native FlashAttention-2 (forward + deterministic backward) behind the torch-facing API of
17ex/flash_attention_dlrs.  Hot path: hand-written sm_100a CUDA in csrc/, reached through the C ABI in
include/fa_b200.h (libfa_b200.so).  No Triton, no autotune, no CPU fallback."""
from ._native import AttentionMask  # noqa: F401
from .flash_attention_torch import (  # noqa: F401
    FlashAttention, FlashAttentionDeterministic, convert_triton_dtype, flash_attention)
from .flash_attention_wrappers import flash_attention_backward, flash_attention_forward  # noqa: F401
from .host_pipeline import HostAttentionPipeline, attention_from_host  # noqa: F401
from .ring import RingAttention, ring_attention_backward, ring_attention_forward  # noqa: F401
from .sharding import PeerGatherBuffer, head_range, head_sharded_attention  # noqa: F401

__all__ = [
    "FlashAttention", "FlashAttentionDeterministic", "convert_triton_dtype", "flash_attention",
    "flash_attention_forward", "flash_attention_backward", "head_range", "head_sharded_attention",
    "HostAttentionPipeline", "attention_from_host", "PeerGatherBuffer",
    "RingAttention", "ring_attention_forward", "ring_attention_backward", "AttentionMask",
]
