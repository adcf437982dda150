"""Forward / dK/dV / dQ time against the block density of a sliding-window mask (C3 shape, non-causal)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flash_attention_dlrs_b200 import AttentionMask, _native  # noqa: E402

B, H, N, D = 2, 32, 8192, 128
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(42)
Q, K, V, dO = (torch.randn(B, H, N, D, generator=g).to(torch.bfloat16).to(dev) for _ in range(4))
sc = D ** -0.5


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


i = torch.arange(N, device=dev)
for w in (0, 128, 512, 1024, 2048, 4096, 8192):
    am = AttentionMask.sliding_window(N, w, w, device=dev) if os.environ.get("BAND") else AttentionMask((i[:, None] - i[None, :]).abs() <= w)
    kw = dict(attn_mask=am)
    O, L = _native.forward(Q, K, V, False, sc, **kw)
    delta = _native.backward_preprocess(O, dO)
    f = t(lambda: _native.forward(Q, K, V, False, sc, **kw))
    dkdv = t(lambda: _native.backward(Q, K, V, O, dO, L, False, sc, 1, delta, **kw))
    dq = t(lambda: _native.backward(Q, K, V, O, dO, L, False, sc, 2, delta, **kw))
    print(f"window +-{w}: block density {(am.blocks > 0).float().mean().item():.3f}  fwd {f:.3f} ms  dkdv {dkdv:.3f}  dq {dq:.3f}")
