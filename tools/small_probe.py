"""Host-side launch overhead vs GPU time at small shapes (where the step is launch-bound)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flash_attention_dlrs_b200 import _native, FlashAttention
dev = torch.device("cuda", 0)
for (B, H, N, D) in [(8, 16, 512, 128), (8, 16, 1024, 128), (8, 16, 512, 64)]:
    g = torch.Generator().manual_seed(0)
    Q, K, V, dO = (torch.randn(B, H, N, D, generator=g).to(torch.float16).to(dev) for _ in range(4))
    def step():
        O, L = _native.forward(Q, K, V, False, 1.0)
        return _native.backward(Q, K, V, O, dO, L, False, 1.0)
    for _ in range(20): step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200): step()
    host = (time.perf_counter() - t0) / 200
    torch.cuda.synchronize()
    total = (time.perf_counter() - t0) / 200
    # GPU-only time via graph replay
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step()
    torch.cuda.current_stream().wait_stream(s)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        out = step()
    gr.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(200): gr.replay()
    b.record(); torch.cuda.synchronize()
    q, k, v = (t.clone().requires_grad_(True) for t in (Q, K, V))
    def sdpa():
        with torch.nn.attention.sdpa_kernel(torch.nn.attention.SDPBackend.FLASH_ATTENTION):
            o = torch.nn.functional.scaled_dot_product_attention(q, k, v, scale=1.0)
        o.backward(dO)
    for _ in range(20): sdpa()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200): sdpa()
    torch.cuda.synchronize()
    ref = (time.perf_counter() - t0) / 200
    print("B%d H%d N%d D%d  host issue %.1f us/step, wall %.1f us/step, GPU-only (graph) %.1f us/step | torch SDPA-flash fwd+bwd wall %.1f us" % (
        B, H, N, D, host * 1e6, total * 1e6, a.elapsed_time(b) / 200 * 1e3, ref * 1e6), flush=True)
