"""How do the kernels behave under sustained load (clocks / power)?  Times blocks of launches while nvidia-smi samples."""
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
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown", "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [lines.append((time.time(), l.strip())) for l in proc.stdout], daemon=True).start()
def block(name, fn, blocks=8, reps=100):
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
        smp = [l for (t, l) in lines if t0 <= t <= t1]
        clk = [float(s.split(",")[0]) for s in smp if s]
        pw = [float(s.split(",")[1]) for s in smp if s]
        cap = any("Active" == s.split(",")[3].strip() for s in smp if s)
        out.append("%.3fms@%s/%sW%s" % (a.elapsed_time(b) / reps, int(sum(clk) / len(clk)) if clk else "?", int(sum(pw) / len(pw)) if pw else "?", "*" if cap else ""))
    print(name, " ".join(out), flush=True)
block("fwd ", lambda: _native.forward(Q, K, V, True, sc))
block("dkdv", lambda: _native.backward(Q, K, V, O, dO, L, True, sc, 1, delta))
block("dq  ", lambda: _native.backward(Q, K, V, O, dO, L, True, sc, 2, delta))
def step():
    o, l = _native.forward(Q, K, V, True, sc)
    _native.backward(Q, K, V, o, dO, l, True, sc)
block("step", step, blocks=8, reps=40)
a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16); b2 = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
block("gemm", lambda: torch.matmul(a, b2), blocks=6, reps=100)
proc.terminate()
