"""Cost of the in-kernel dropout on the C3 shape (bf16 B2 H32 N8192 D128 causal): forward, dK/dV and dQ kernels with
dropout_p = 0 and 0.1, CUDA events, 20 launches each after 5 warm-ups."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flash_attention_dlrs_b200 import _native  # noqa: E402

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


flops = 4.0 * B * H * N * N * D * 0.5
for p in (0.0, 0.1):
    kw = dict(dropout_p=p, dropout_seed=7)
    O, L = _native.forward(Q, K, V, True, sc, **kw)
    delta = _native.backward_preprocess(O, dO)
    f = t(lambda: _native.forward(Q, K, V, True, sc, **kw))
    dkdv = t(lambda: _native.backward(Q, K, V, O, dO, L, True, sc, 1, delta, **kw))
    dq = t(lambda: _native.backward(Q, K, V, O, dO, L, True, sc, 2, delta, **kw))
    print(f"dropout_p {p}: fwd {f:.3f} ms ({flops / f / 1e9:.0f} TFLOP/s)  dkdv {dkdv:.3f}  dq {dq:.3f}  "
          f"fwd+bwd {flops * 3.5 / (f + dkdv + dq) / 1e9:.0f} TFLOP/s (kernels only)")
