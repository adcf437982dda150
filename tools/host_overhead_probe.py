"""Where does the host time of a small fwd+bwd step go?  cProfile over the autograd path at a launch-bound shape."""
import cProfile, os, pstats, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flash_attention_dlrs_b200 import FlashAttention, _native
dev = torch.device("cuda", 0)
B, H, N, D = 8, 16, 512, 64
g = torch.Generator().manual_seed(0)
Q, K, V, dO = (torch.randn(B, H, N, D, generator=g).to(torch.float16).to(dev) for _ in range(4))
q, k, v = (t.clone().requires_grad_(True) for t in (Q, K, V))
def step():
    q.grad = k.grad = v.grad = None
    FlashAttention.apply(q, k, v, False, 1.0).backward(dO)
def native():
    O, L = _native.forward(Q, K, V, False, 1.0)
    _native.backward(Q, K, V, O, dO, L, False, 1.0)
for fn, name in ((step, "autograd"), (native, "native")):
    for _ in range(50): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(500): fn()
    host = (time.perf_counter() - t0) / 500
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 500
    print(f"{name}: host issue {host*1e6:.1f} us/step, wall {wall*1e6:.1f} us/step", flush=True)
pr = cProfile.Profile()
pr.enable()
for _ in range(500): native()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
