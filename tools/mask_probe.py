"""Cost of an arbitrary attention mask (one (N,N) byte mask shared by all heads) on the C3 shape, non-causal:
forward, dK/dV and dQ kernels with and without the mask, CUDA events, 20 launches each after 5 warm-ups."""
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


def t(fn, reps=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


flops = 4.0 * B * H * N * N * D
masks = {"none": None,
         "random 90% (N,N)": AttentionMask(torch.rand(N, N, device=dev) < 0.9),
         "random 90% (B,H,N,N)": AttentionMask(torch.rand(B, H, N, N, device=dev) < 0.9),
         "sliding window +-512 (N,N)": AttentionMask((torch.arange(N, device=dev)[:, None]
                                                      - torch.arange(N, device=dev)[None, :]).abs() <= 512),
         "block-sparse 25% (N,N)": AttentionMask((torch.rand(N // 128, N // 128, device=dev) < 0.25)
                                                 .repeat_interleave(128, 0).repeat_interleave(128, 1))}
for name, am in masks.items():
    kw = dict(attn_mask=am)
    O, L = _native.forward(Q, K, V, False, sc, **kw)
    delta = _native.backward_preprocess(O, dO)
    f = t(lambda: _native.forward(Q, K, V, False, sc, **kw))
    dkdv = t(lambda: _native.backward(Q, K, V, O, dO, L, False, sc, 1, delta, **kw))
    dq = t(lambda: _native.backward(Q, K, V, O, dO, L, False, sc, 2, delta, **kw))
    dens = 1.0 if am is None else (am.blocks > 0).float().mean().item()
    print(f"mask {name} (block density {dens:.2f}): fwd {f:.3f} ms ({flops / f / 1e9:.0f} TFLOP/s)  dkdv {dkdv:.3f}  dq {dq:.3f}  "
          f"fwd+bwd {flops * 3.5 / (f + dkdv + dq) / 1e9:.0f} TFLOP/s (dense-equivalent, kernels only)")
