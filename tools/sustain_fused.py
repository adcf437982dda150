"""Sustained (power-capped) time per matmul unit: two-kernel backward vs the single-pass pipeline without its dQ egress."""
import os, subprocess, sys, time, threading, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flash_attention_dlrs_b200 import _native
B, H, N, D = 2, 32, 8192, 128
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(42)
Q, K, V, dO = (torch.randn(B, H, N, D, generator=g).to(torch.bfloat16).to(dev) for _ in range(4))
sc = D ** -0.5
O, L = _native.forward(Q, K, V, True, sc)
delta = _native.backward_preprocess(O, dO)
lines = []
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [lines.append((time.time(), l.strip())) for l in proc.stdout], daemon=True).start()
def block(name, fn, units, blocks=6, reps=100):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(blocks):
        t0 = time.time()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): fn()
        b.record(); torch.cuda.synchronize()
        t1 = time.time()
        smp = [l for (t, l) in lines if t0 <= t <= t1 and l]
        clk = [float(s.split(",")[0]) for s in smp]; pw = [float(s.split(",")[1]) for s in smp]
        ms = a.elapsed_time(b) / reps
        out.append("%.3fms(%.3f/unit)@%s/%sW" % (ms, ms / units, int(sum(clk) / len(clk)) if clk else "?", int(sum(pw) / len(pw)) if pw else "?"))
    print(name, " ".join(out), flush=True)
block("dkdv  (4 units)", lambda: _native.backward(Q, K, V, O, dO, L, True, sc, 1, delta), 4)
block("dq    (3 units)", lambda: _native.backward(Q, K, V, O, dO, L, True, sc, 2, delta), 3)
block("fused (5 units)", lambda: _native.backward(Q, K, V, O, dO, L, True, sc, 4, delta), 5)
block("fwd   (2 units)", lambda: _native.forward(Q, K, V, True, sc), 2)
proc.terminate()
