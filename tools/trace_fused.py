"""Schedule of the whole fused-backward grid (debug build with -DFA_TRACE=2): per-CTA globaltimer milestones."""
import ctypes, os, sys, torch
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("FA_B200_LIB", os.path.join(root, "build", "libfa_trace2.so"))
sys.path.insert(0, root)
from flash_attention_dlrs_b200 import _lib, _native
lib = _lib.load()
B, H, N, D = 2, 32, 8192, 128
causal = (sys.argv[1] != "0") if len(sys.argv) > 1 else True
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(42)
Q, K, V, dO = (torch.randn(B, H, N, D, generator=g).to(torch.bfloat16).to(dev) for _ in range(4))
sc = D ** -0.5
O, L = _native.forward(Q, K, V, causal, sc)
delta = _native.backward_preprocess(O, dO)
run = lambda: _native.backward(Q, K, V, O, dO, L, causal, sc, _native.BWD_FUSED, delta)
run(); torch.cuda.synchronize()
n_cta = B * H * (N // 128)
buf = torch.zeros(n_cta * 8, dtype=torch.int64, device=dev)
lib.fa_debug_set_trace(ctypes.c_void_p(buf.data_ptr()), n_cta * 8)
run(); torch.cuda.synchronize()
ev = buf.cpu().view(n_cta, 8).double()
t0 = ev[:, 0].min()
entry, setup, firstp, loopend, exit_, smid, nvis = [ev[:, k] for k in range(7)]
print("grid span %.1f us, CTAs %d" % ((exit_.max() - t0) / 1e3, n_cta))
print("set-up (entry->sync)      mean %.2f us  max %.2f" % ((setup - entry).mean() / 1e3, (setup - entry).max() / 1e3))
print("fill   (sync->first P)    mean %.2f us  max %.2f" % ((firstp - setup).mean() / 1e3, (firstp - setup).max() / 1e3))
per_visit = (loopend - firstp) / nvis
print("loop   per visit          mean %.2f us  (heavy CTAs: %.2f, light: %.2f)" % (per_visit.mean() / 1e3, per_visit[nvis > 32].mean() / 1e3, per_visit[nvis <= 8].mean() / 1e3))
print("drain  (loop end->exit)   mean %.2f us  max %.2f" % ((exit_ - loopend).mean() / 1e3, (exit_ - loopend).max() / 1e3))
# gap between a CTA's exit and the next CTA's entry on the same SM
gaps = []
for sm in smid.unique():
    idx = (smid == sm).nonzero().flatten()
    o = idx[entry[idx].argsort()]
    gaps.append(entry[o][1:] - exit_[o][:-1])
gaps = torch.cat(gaps)
print("gap    (exit->next entry) mean %.2f us  max %.2f  SMs used %d" % (gaps.mean() / 1e3, gaps.max() / 1e3, len(smid.unique())))
busy = (exit_ - entry).sum() / 1e3
print("sum of CTA lifetimes %.0f us -> %.1f us per SM; visits total %d -> %.3f us per visit overall" % (busy, busy / 148, int(nvis.sum()), (exit_.max() - t0) / 1e3 * 148 / nvis.sum()))
last = exit_.argsort()[-5:]
print("last CTAs to finish: tickets", last.tolist(), "visits", nvis[last].tolist(), "entry(us)", ((entry[last] - t0) / 1e3).tolist())
first_idle = torch.stack([exit_[smid == sm].max() for sm in smid.unique()])
print("per-SM finish time: min %.1f  mean %.1f  max %.1f us" % ((first_idle.min() - t0) / 1e3, (first_idle.mean() - t0) / 1e3, (first_idle.max() - t0) / 1e3))
