"""Forward timing + accuracy for alternative builds (FA_B200_LIB): config 3 causal, config 2 non-causal D=64."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flash_attention_dlrs_b200 import _native
dev = torch.device("cuda", 0)
def t(fn, reps=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for (B, H, N, D, causal, dt) in [(2, 32, 8192, 128, True, torch.bfloat16), (2, 32, 8192, 128, False, torch.bfloat16), (4, 16, 4096, 64, False, torch.float16),
                                 (2, 32, 8192, 128, True, torch.float8_e4m3fn), (2, 32, 8192, 128, False, torch.float8_e4m3fn), (2, 32, 8192, 128, True, torch.float8_e5m2)]:
    g = torch.Generator().manual_seed(42)
    Q, K, V = (torch.randn(B, H, N, D, generator=g).to(dt).to(dev) for _ in range(3))
    sc = D ** -0.5
    ms = t(lambda: _native.forward(Q, K, V, causal, sc))
    fl = 4.0 * B * H * N * N * D * (0.5 if causal else 1.0)
    O, L = _native.forward(Q[:, :2, :1024], K[:, :2, :1024], V[:, :2, :1024], causal, sc)
    ref = torch.nn.functional.scaled_dot_product_attention(Q[:, :2, :1024].float(), K[:, :2, :1024].float(), V[:, :2, :1024].float(), is_causal=causal, scale=sc)
    err = (O.float() - ref).abs().max().item()
    print("lib %s %s B%d H%d N%d D%d causal=%d: %.3f ms  %.0f TFLOP/s   max|O-ref| %.2e" % (os.path.basename(os.environ.get("FA_B200_LIB", "default")), str(dt).split(".")[-1], B, H, N, D, causal, ms, fl / ms / 1e9, err), flush=True)
