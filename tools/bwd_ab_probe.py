"""Backward kernels A/B between two builds (FA_B200_LIB): times of dK/dV and dQ on a few shapes + checksums of dQ/dK/dV."""
import hashlib, os, sys, torch
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
shapes = [(2, 3, n, d, c, dt) for d in (128, 64) for n in (1, 100, 128, 200, 384, 640, 1100) for c in (False, True) for dt in (torch.bfloat16, torch.float16)]
shapes += [(2, 32, 8192, 128, True, torch.bfloat16), (2, 32, 8192, 128, False, torch.bfloat16), (4, 16, 4096, 64, False, torch.float16),
           (8, 16, 512, 128, False, torch.float16), (8, 16, 8192, 64, True, torch.float16), (1, 16, 32768, 128, True, torch.bfloat16)]
for (B, H, N, D, causal, dt) in shapes:
    g = torch.Generator().manual_seed(N + D)
    Q, K, V, dO = (torch.randn(B, H, N, D, generator=g).to(dt).to(dev) for _ in range(4))
    sc = D ** -0.5
    O, L = _native.forward(Q, K, V, causal, sc)
    delta = _native.backward_preprocess(O, dO)
    dQ, dK, dV = _native.backward(Q, K, V, O, dO, L, causal, sc, 3, delta)
    torch.cuda.synchronize()
    hs = hashlib.sha1(b"".join(x.cpu().view(torch.int16).numpy().tobytes() for x in (dQ, dK, dV))).hexdigest()[:10]
    kv = dq = 0.0
    if N >= 512:
        kv = t(lambda: _native.backward(Q, K, V, O, dO, L, causal, sc, 1, delta))
        dq = t(lambda: _native.backward(Q, K, V, O, dO, L, causal, sc, 2, delta))
    print("lib %s B%d H%d N%d D%d c%d %s  %s  dkdv %.3f dq %.3f ms" % (tag, B, H, N, D, causal, str(dt)[6:], hs, kv, dq), flush=True)
