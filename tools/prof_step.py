"""Two fwd+bwd steps of the bench workload (BASELINE configs[2]) for ncu: profile the second step.
`fused` as last argument adds the optional single-pass backward (FA_BWD_FUSED) after each step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from flash_attention_dlrs_b200 import _native

args = [a for a in sys.argv[1:] if a != "fused"]
fused = "fused" in sys.argv[1:]
B, H, N, D = 2, 32, 8192, 128
if len(args) >= 4:
    B, H, N, D = (int(x) for x in args[:4])
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(42)
Q, K, V, dO = (torch.randn(B, H, N, D, generator=g).to(torch.bfloat16).to(dev) for _ in range(4))
scale = D ** -0.5
for _ in range(2):
    O, L = _native.forward(Q, K, V, True, scale)
    _native.backward(Q, K, V, O, dO, L, True, scale)
    if fused:
        _native.backward(Q, K, V, O, dO, L, True, scale, which=_native.BWD_FUSED)
torch.cuda.synchronize()
print("ok")
