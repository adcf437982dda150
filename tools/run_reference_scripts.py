#!/usr/bin/env python
"""Run the reference's OWN test scripts, unmodified, against this library: src/test_torch.py (gradcheck of both autograd
Functions) and src/test_correctness.py (200 seeded fp32 comparisons with torch SDPA at the reference's tolerances).
The scripts are read from baseline/_ref/src (an untouched copy of the reference's sources, placed there by
tools/fetch_reference.sh; git-ignored); only sys.path decides that `flash_attention_torch` / `flash_attention_wrappers`
resolve to flash_attention_dlrs_b200/compat instead of the reference's Triton modules.

    python tools/run_reference_scripts.py [test_torch.py] [test_correctness.py]
"""
import os
import runpy
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "baseline", "_ref", "src")
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash_attention_dlrs_b200", "compat")]

scripts = sys.argv[1:] or ["test_torch.py", "test_correctness.py"]
for name in scripts:
    path = os.path.join(SRC, name)
    if not os.path.exists(path):
        print(f"{name}: {path} not present (run tools/fetch_reference.sh in the dev container)")
        sys.exit(2)
    print(f"=== {name} (unmodified reference script, sha1 of file below) ===", flush=True)
    import hashlib
    print(hashlib.sha1(open(path, "rb").read()).hexdigest(), flush=True)
    t0 = time.time()
    runpy.run_path(path, run_name="__main__")
    import flash_attention_wrappers as w   # prove which module the script imported
    print(f"[{name}: {time.time() - t0:.1f} s; flash_attention_wrappers -> {w.__file__}]", flush=True)
