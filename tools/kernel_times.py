"""Per-kernel times of the default path on the bench workload (C3) and the C2 forward; FA_B200_LIB selects the build."""
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
tag = os.path.basename(os.environ.get("FA_B200_LIB", "default"))
for (B, H, N, D, causal, dt) in [(2, 32, 8192, 128, True, torch.bfloat16), (4, 16, 4096, 64, False, torch.float16)]:
    g = torch.Generator().manual_seed(42)
    Q, K, V, dO = (torch.randn(B, H, N, D, generator=g).to(dt).to(dev) for _ in range(4))
    sc = D ** -0.5
    O, L = _native.forward(Q, K, V, causal, sc)
    delta = _native.backward_preprocess(O, dO)
    f = t(lambda: _native.forward(Q, K, V, causal, sc))
    pre = t(lambda: _native.backward_preprocess(O, dO))
    kv = t(lambda: _native.backward(Q, K, V, O, dO, L, causal, sc, 1, delta))
    dq = t(lambda: _native.backward(Q, K, V, O, dO, L, causal, sc, 2, delta))
    def step():
        o, l = _native.forward(Q, K, V, causal, sc)
        _native.backward(Q, K, V, o, dO, l, causal, sc)
    st = t(step)
    print("lib %-22s B%d H%d N%d D%d c%d  fwd %.3f  pre %.3f  dkdv %.3f  dq %.3f  step %.3f ms" % (tag, B, H, N, D, causal, f, pre, kv, dq, st), flush=True)
