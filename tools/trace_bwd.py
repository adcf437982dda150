"""Timelines of one CTA of the dK/dV and dQ kernels (debug build with -DFA_TRACE=1): clock64 per role / block / event."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flash_attention_dlrs_b200 import _lib, _native
lib = _lib.load()
B, H, N, D = 2, 32, 8192, 128
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(42)
Q, K, V, dO = (torch.randn(B, H, N, D, generator=g).to(torch.bfloat16).to(dev) for _ in range(4))
sc = D ** -0.5
causal = os.environ.get("FA_TRACE_CAUSAL", "0") != "0"
O, L = _native.forward(Q, K, V, causal, sc)
delta = _native.backward_preprocess(O, dO)
for which, setter, names in ((1, "fa_debug_set_trace_dkdv", "mma: [top, P^T a ready, grad a + score a(t+1) issued, P^T b ready, grad b + score b(t+1) issued]; wg: [wait, scores ready, P/dS stored]"),
                             (2, "fa_debug_set_trace_dq", "mma: [top, score a(t+1) issued, scores(t+1) issued, dS a ready, dQ a issued, dQ b issued]; wg: [wait, scores ready, dS handed over]")):
    run = lambda: _native.backward(Q, K, V, O, dO, L, causal, sc, which, delta)
    run(); torch.cuda.synchronize()
    roles = 4
    buf = torch.zeros(roles * 8192, dtype=torch.int64, device=dev)
    getattr(lib, setter)(ctypes.c_void_p(buf.data_ptr()), roles * 8192)
    run(); torch.cuda.synchronize()
    getattr(lib, setter)(ctypes.c_void_p(0), 0)
    ev = buf.cpu().view(roles, 1024, 8)
    print("== kernel", "dK/dV" if which == 1 else "dQ", "|", names)
    for it in (10, 11, 12, 30, 31):
        base = ev[0, it, 0].item()
        print("it", it, "| period", ev[0, it + 1, 0].item() - base, "| mma", [int(x) - base for x in ev[0, it].tolist() if x > 0],
              "| wg_a", [int(x) - base for x in ev[1, it].tolist() if x > 0], "| wg_b", [int(x) - base for x in ev[2, it].tolist() if x > 0])
