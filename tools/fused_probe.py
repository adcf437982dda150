"""Timing of the single-pass backward (fa_bwd, FA_BWD_FUSED) against the two-kernel path; FA_B200_LIB selects the build."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flash_attention_dlrs_b200 import _native
dev = torch.device("cuda", 0)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
cases = [(2, 32, 8192, 128, True), (2, 32, 8192, 128, False), (4, 16, 4096, 64, False), (1, 8, 32768, 128, True)]
if len(sys.argv) > 1: cases = cases[:int(sys.argv[1])]
for (B, H, N, D, causal) in cases:
    g = torch.Generator().manual_seed(42)
    Q, K, V, dO = (torch.randn(B, H, N, D, generator=g).to(torch.bfloat16).to(dev) for _ in range(4))
    sc = D ** -0.5
    O, L = _native.forward(Q, K, V, causal, sc)
    delta = _native.backward_preprocess(O, dO)
    unit = 2.0 * B * H * N * N * D * (0.5 if causal else 1.0)
    f = t(lambda: _native.backward(Q, K, V, O, dO, L, causal, sc, _native.BWD_FUSED, delta))
    two = t(lambda: _native.backward(Q, K, V, O, dO, L, causal, sc, 3, delta))
    print("lib %s  B%d H%d N%d D%d causal=%d  fused %.3f ms (%.0f TFLOP/s alg)  two-kernel %.3f ms (%.0f)" % (
        os.path.basename(os.environ.get("FA_B200_LIB", "default")), B, H, N, D, causal, f, 5 * unit / f / 1e9, two, 5 * unit / two / 1e9), flush=True)
