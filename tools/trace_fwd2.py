"""Timeline of the heaviest CTA pair of the pair forward (debug build with -DFA_TRACE=1): clock64 per role / block / event."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("FA_FWD_PAIR", "1")
pair = os.environ["FA_FWD_PAIR"] != "0"
from flash_attention_dlrs_b200 import _lib, _native
lib = _lib.load()
B, H, N, D = 2, 32, 8192, 128
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
names = {0: "mma", 1: "L.sm0", 2: "L.sm1", 4: "F.sm0", 5: "F.sm1"}
# mma: [top0, P0a, P0b, S0(j+1) issued, top1, P1a, P1b, S1(j+1) issued]; sm: [wait S, S ready, P half 1 arrived, P half 2 arrived]
for it in (10, 11, 12, 30, 31, 50):
    base = ev[0, it, 0].item()
    print("it", it, "| mma period", ev[0, it + 1, 0].item() - base, "| mma", [int(x) - base for x in ev[0, it].tolist()],
          "| leader sm0", [int(x) - base for x in ev[1, it, :4].tolist()], "sm1", [int(x) - base for x in ev[2, it, :4].tolist()])
    if not pair:
        continue
    fb = ev[4, it, 0].item()
    print("      follower (own clock) sm0", [int(x) - fb for x in ev[4, it, :4].tolist()], "period", ev[4, it + 1, 0].item() - fb,
          "sm1", [int(x) - fb for x in ev[5, it, :4].tolist()])
