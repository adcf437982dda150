"""Timeline of the heaviest CTA of the duo forward (debug build with -DFA_TRACE=1 -DFA_EXPERIMENTAL_FWD=1, FA_FWD_DUO=1)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("FA_FWD_DUO", "1")
from flash_attention_dlrs_b200 import _lib, _native
lib = _lib.load()
B, H, N, D = 2, 32, 8192, int(os.environ.get("FA_TRACE_D", "128"))
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(42)
Q, K, V = (torch.randn(B, H, N, D, generator=g).to(torch.bfloat16).to(dev) for _ in range(3))
sc = D ** -0.5
causal = os.environ.get("FA_TRACE_CAUSAL", "0") != "0"
run = lambda: _native.forward(Q, K, V, causal, sc)
run(); torch.cuda.synchronize()
roles = 6
buf = torch.zeros(roles * 8192, dtype=torch.int64, device=dev)
lib.fa_debug_set_trace_fwd.restype = ctypes.c_int
lib.fa_debug_set_trace_fwd(ctypes.c_void_p(buf.data_ptr()), roles * 8192)
run(); torch.cuda.synchronize()
ev = buf.cpu().view(roles, 1024, 8)
# mma: [top0, P0a, P0b, S0(j+1) issued, top1, P1a, P1b, S1(j+1) issued]
# softmax warp (role 1 + 2 hf + t): [wait S, S ready, before pair barrier, after it, first chunk of P arrived, second chunk arrived]
for it in (10, 11, 12, 30, 31, 50):
    base = ev[0, it, 0].item()
    print("it", it, "| mma period", ev[0, it + 1, 0].item() - base, "| mma", [int(x) - base for x in ev[0, it].tolist()])
    for r, nm in ((1, "w0 t0"), (3, "w4 t0"), (2, "w0 t1"), (4, "w4 t1")):
        print("      %s" % nm, [int(x) - base for x in ev[r, it, :6].tolist()])
