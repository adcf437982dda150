"""Flat-module shim for `from flash_attention_wrappers import flash_attention_forward, flash_attention_backward`
(src/test_correctness.py:1)."""
from flash_attention_dlrs_b200.flash_attention_wrappers import (  # noqa: F401
    flash_attention_backward, flash_attention_forward)
