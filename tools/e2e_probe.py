"""PCIe behaviour of the host pipeline on this box: sequential copies, simplex pipeline, duplex pipeline."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flash_attention_dlrs_b200 import HostAttentionPipeline
B, H, N, D = 2, 32, 8192, 128
dev = torch.device("cuda", 0)
host = [torch.randn(B, H, N, D).to(torch.bfloat16).pin_memory() for _ in range(4)]
out = [torch.empty(B, H, N, D, dtype=torch.bfloat16).pin_memory() for _ in range(4)]
dev_t = [t.to(dev) for t in host]
def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
h2d = timeit(lambda: [d.copy_(h, non_blocking=True) for d, h in zip(dev_t, host)])
d2h = timeit(lambda: [h.copy_(d, non_blocking=True) for d, h in zip(dev_t, out)])
print("H2D 512MiB %.2f ms = %.1f GB/s ; D2H %.2f ms = %.1f GB/s" % (h2d, 0.537 / h2d * 1e3, d2h, 0.537 / d2h * 1e3))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        for d, h in zip(dev_t, host): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        for d, h in zip(dev_t, out): h.copy_(d, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
bd = timeit(both)
print("simultaneous H2D + D2H %.2f ms = %.1f GB/s each way" % (bd, 0.537 / bd * 1e3))
for duplex in (False, True):
    for chunks in (4, 8, 16, 32):
        pipe = HostAttentionPipeline(B, H, N, D, torch.bfloat16, dev, chunks=chunks, duplex=duplex)
        def steps(n=5):   # n pipelined steps; the stream waits for the last download before the closing event
            for _ in range(n):
                done = pipe.run(host, out, True, D ** -0.5)
            torch.cuda.current_stream().wait_event(done)
        ms = timeit(steps, reps=2) / 5
        print("pipeline duplex=%s chunks=%d: %.2f ms" % (duplex, chunks, ms))
