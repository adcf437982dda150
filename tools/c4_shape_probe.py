import os, sys, torch
sys.path.insert(0, os.getcwd())
from flash_attention_dlrs_b200 import _native, flash_attention_forward, flash_attention_backward
dev = torch.device("cuda", 0)
def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for (B, H, N, D) in [(1, 8, 32768, 128), (1, 32, 32768, 128), (2, 32, 8192, 128)]:
    g = torch.Generator().manual_seed(42)
    Q, K, V, dO = (torch.randn(B, H, N, D, generator=g).to(torch.bfloat16).to(dev) for _ in range(4))
    sc = D ** -0.5
    O, L = _native.forward(Q, K, V, True, sc)
    delta = _native.backward_preprocess(O, dO)
    f = t(lambda: _native.forward(Q, K, V, True, sc))
    kv = t(lambda: _native.backward(Q, K, V, O, dO, L, True, sc, 1, delta))
    dq = t(lambda: _native.backward(Q, K, V, O, dO, L, True, sc, 2, delta))
    def step():
        o, l = flash_attention_forward(Q, K, V, dev, True, sc)
        return flash_attention_backward(Q, K, V, o, dO, l, dev, True, True, sc)
    st = t(step, 3)
    print("B%d H%d N%d: fwd %.3f dkdv %.3f dq %.3f step %.3f ms" % (B, H, N, f, kv, dq, st), flush=True)
