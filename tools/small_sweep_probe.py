"""Short sequences: GPU-only time per kernel (CUDA graph replay of each launch) next to the host time of the autograd step, for
the C5 shapes N = 128 .. 1024 — which part of a launch-bound step is the kernels' own latency and which is the host path."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flash_attention_dlrs_b200 import _native, FlashAttention
from flash_attention_dlrs_b200 import flash_attention_torch as fat
dev = torch.device("cuda", 0)
# FA_PROBE_PY=1: the Python autograd Function instead of the C++ node (csrc/torch_binding.cpp)
fat.USE_CPP_NODE = os.environ.get("FA_PROBE_PY") != "1"
print("autograd node:", "C++" if fat.USE_CPP_NODE and fat._cpp_node() is not None else "Python", flush=True)

def graph_us(fn, reps=200):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        keep = fn()
    gr.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): gr.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3

for D in (64, 128):
    for causal in (False, True):
        for N in (128, 512, 1024):
            B, H = 8, 16
            g = torch.Generator().manual_seed(0)
            Q, K, V, dO = (torch.randn(B, H, N, D, generator=g).to(torch.float16).to(dev) for _ in range(4))
            O, L = _native.forward(Q, K, V, causal, 1.0)
            delta = _native.backward_preprocess(O, dO)
            f = graph_us(lambda: _native.forward(Q, K, V, causal, 1.0))
            p = graph_us(lambda: _native.backward_preprocess(O, dO))
            kv = graph_us(lambda: _native.backward(Q, K, V, O, dO, L, causal, 1.0, 1, delta))
            dq = graph_us(lambda: _native.backward(Q, K, V, O, dO, L, causal, 1.0, 2, delta))
            q, k, v = (t.clone().requires_grad_(True) for t in (Q, K, V))
            def step():
                q.grad = k.grad = v.grad = None
                FlashAttention.apply(q, k, v, causal, 1.0).backward(dO)
            for _ in range(30): step()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(300): step()
            host = (time.perf_counter() - t0) / 300
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) / 300
            def fwd_only():
                with torch.no_grad():
                    FlashAttention.apply(Q, K, V, causal, 1.0)
            for _ in range(30): fwd_only()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(300): fwd_only()
            torch.cuda.synchronize()
            fwall = (time.perf_counter() - t0) / 300
            print("D%d c%d N%4d | GPU us: fwd %5.1f pre %4.1f dkdv %5.1f dq %5.1f sum %6.1f | autograd step: host %6.1f wall %6.1f us | fwd call wall %5.1f us"
                  % (D, causal, N, f, p, kv, dq, f + p + kv + dq, host * 1e6, wall * 1e6, fwall * 1e6), flush=True)
