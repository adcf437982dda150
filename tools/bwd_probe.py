"""Two-kernel backward timing + accuracy for alternative builds (FA_B200_LIB)."""
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
for (B, H, N, D, causal, dt) in [(2, 32, 8192, 128, True, torch.bfloat16), (2, 32, 8192, 128, False, torch.bfloat16), (4, 16, 4096, 64, False, torch.float16)]:
    g = torch.Generator().manual_seed(42)
    Q, K, V, dO = (torch.randn(B, H, N, D, generator=g).to(dt).to(dev) for _ in range(4))
    sc = D ** -0.5
    O, L = _native.forward(Q, K, V, causal, sc)
    delta = _native.backward_preprocess(O, dO)
    a = t(lambda: _native.backward(Q, K, V, O, dO, L, causal, sc, 1, delta))
    b = t(lambda: _native.backward(Q, K, V, O, dO, L, causal, sc, 2, delta))
    # accuracy on a slice against fp32 autograd
    q, k, v = (x[:, :2, :1024].float().requires_grad_(True) for x in (Q, K, V))
    o = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=causal, scale=sc)
    gq, gk, gv = torch.autograd.grad(o, (q, k, v), dO[:, :2, :1024].float())
    Os, Ls = _native.forward(Q[:, :2, :1024], K[:, :2, :1024], V[:, :2, :1024], causal, sc)
    dq, dk, dv = _native.backward(Q[:, :2, :1024], K[:, :2, :1024], V[:, :2, :1024], Os, dO[:, :2, :1024], Ls, causal, sc)
    rel = lambda x, r: ((x.float() - r).abs().max() / r.abs().max()).item()
    print("lib %s B%d H%d N%d D%d causal=%d: dkdv %.3f ms  dq %.3f ms  sum %.3f | rel err dq %.2e dk %.2e dv %.2e" % (
        os.path.basename(os.environ.get("FA_B200_LIB", "default")), B, H, N, D, causal, a, b, a + b, rel(dq, gq), rel(dk, gk), rel(dv, gv)), flush=True)
