"""One warm + one profiled pass of the backward kernels on BASELINE config 2 (fp16 B4 H16 N4096 D64 non-causal), for ncu."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flash_attention_dlrs_b200 import _native
dev = torch.device("cuda", 0)
B, H, N, D = 4, 16, 4096, 64
g = torch.Generator().manual_seed(42)
Q, K, V, dO = (torch.randn(B, H, N, D, generator=g).to(torch.float16).to(dev) for _ in range(4))
sc = D ** -0.5
for _ in range(2):
    O, L = _native.forward(Q, K, V, False, sc)
    delta = _native.backward_preprocess(O, dO)
    _native.backward(Q, K, V, O, dO, L, False, sc, 1, delta)
    _native.backward(Q, K, V, O, dO, L, False, sc, 2, delta)
    torch.cuda.synchronize()
print("ok")
