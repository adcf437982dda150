"""Timeline of one CTA of a backward kernel (debug build with -DFA_TRACE=1): clock64 per role / iteration / event."""
import ctypes, os, sys, torch
os.environ.setdefault("FA_B200_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "build", "libfa_trace.so"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flash_attention_dlrs_b200 import _lib, _native
lib = _lib.load()
B, H, N, D = 2, 32, 8192, 128
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(42)
Q, K, V, dO = (torch.randn(B, H, N, D, generator=g).to(torch.bfloat16).to(dev) for _ in range(4))
sc = D ** -0.5
O, L = _native.forward(Q, K, V, True, sc)
delta = _native.backward_preprocess(O, dO)
which = int(sys.argv[1]) if len(sys.argv) > 1 else 2
causal = os.environ.get("FA_TRACE_CAUSAL", "1") != "0"
O, L = _native.forward(Q, K, V, causal, sc)
delta = _native.backward_preprocess(O, dO)
run = (lambda: _native.forward(Q, K, V, causal, sc)) if which == 0 else (lambda: _native.backward(Q, K, V, O, dO, L, causal, sc, which, delta))
run(); torch.cuda.synchronize()
roles = 4
buf = torch.zeros(roles * 8192, dtype=torch.int64, device=dev)
lib.fa_debug_set_trace(ctypes.c_void_p(buf.data_ptr()), roles * 8192)
run(); torch.cuda.synchronize()
ev = buf.cpu().view(roles, 1024, 8)
t0 = ev[ev > 0].min().item()
names = {0: "mma", 1: "wg_a", 2: "wg_b", 3: "rd/prod"}
for it in (10, 11, 12, 30, 31):
    base = ev[0, it, 0].item()
    print("it", it, "| period", ev[0, it + 1, 0].item() - base, "|",
          {names[r]: [int(x) - base for x in ev[r, it].tolist() if x > 0] for r in range(roles)})
