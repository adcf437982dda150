"""Host<->device copy bandwidth on this box: direction, simultaneity, NUMA placement of the pinned buffer, copy size."""
import os
import subprocess
import sys

import torch

print(subprocess.run("lscpu | grep -i -E 'numa|model name|^CPU\\(s\\)'; nvidia-smi topo -m | head -8", shell=True,
                     capture_output=True, text=True).stdout)
dev = torch.device("cuda", 0)
NBYTES = 512 << 20


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def run(tag):
    h_in = torch.empty(NBYTES, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    h_out = torch.empty(NBYTES, dtype=torch.uint8).pin_memory()
    h_out.fill_(2)
    d_in = torch.empty(NBYTES, dtype=torch.uint8, device=dev)
    d_out = torch.ones(NBYTES, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    gb = NBYTES / 1e9
    h2d = timeit(lambda: d_in.copy_(h_in, non_blocking=True))
    d2h = timeit(lambda: h_out.copy_(d_out, non_blocking=True))

    def both(chunk):
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur)
        s2.wait_stream(cur)
        for o in range(0, NBYTES, chunk):
            with torch.cuda.stream(s1):
                d_in[o:o + chunk].copy_(h_in[o:o + chunk], non_blocking=True)
            with torch.cuda.stream(s2):
                h_out[o:o + chunk].copy_(d_out[o:o + chunk], non_blocking=True)
        cur.wait_stream(s1)
        cur.wait_stream(s2)

    line = f"{tag}: H2D {gb / h2d * 1e3:.1f} GB/s  D2H {gb / d2h * 1e3:.1f} GB/s  duplex (each way)"
    for chunk in (NBYTES, 64 << 20, 8 << 20, 1 << 20):
        ms = timeit(lambda: both(chunk))
        line += f"  {chunk >> 20}MiB:{gb / ms * 1e3:.1f}"
    print(line, flush=True)


run("default affinity")
ncpu = os.cpu_count()
all_cpus = sorted(os.sched_getaffinity(0))
print("affinity", all_cpus)
if len(all_cpus) > 1:
    half = len(all_cpus) // 2
    for name, cpus in (("first half", all_cpus[:half]), ("second half", all_cpus[half:])):
        os.sched_setaffinity(0, cpus)
        run(f"pinned while bound to {name} {cpus[0]}-{cpus[-1]}")
    os.sched_setaffinity(0, all_cpus)
