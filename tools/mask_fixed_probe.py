"""Fixed per-CTA cost of the masked forward: block density 0 / one block per tile, several N (grid sizes)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flash_attention_dlrs_b200 import AttentionMask, _native  # noqa: E402

dev = torch.device("cuda", 0)


def t(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for (B, H, N, D) in ((2, 32, 8192, 128), (2, 32, 2048, 128), (8, 32, 2048, 128), (2, 32, 8192, 64)):
    Q, K, V = (torch.randn(B, H, N, D, device=dev).to(torch.bfloat16) for _ in range(3))
    i = torch.arange(N, device=dev)
    for name, m in (("empty", torch.zeros(N, N, dtype=torch.bool, device=dev)), ("diag", (i[:, None] == i[None, :]))):
        am = AttentionMask(m)
        f = t(lambda: _native.forward(Q, K, V, False, D ** -0.5, attn_mask=am))
        am.blocks = None
        am._struct = None
        print(f"B{B} H{H} N{N} D{D} mask {name}: fwd {f * 1e3:.0f} us, CTAs {B * H * ((N + 255) // 256)}, "
              f"per wave {f * 1e3 / (B * H * ((N + 255) // 256) / 148):.1f} us")
