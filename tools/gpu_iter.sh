#!/bin/bash
# fast iteration: 16-bit parity subset + bench without the CPU leg
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x --timeout 300 -k "${1:-16bit or config2 or config3 or bit_identical or head_slice or autograd or two_kernel}" > gpurun_out/pytest_iter.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_iter.log; tail -15 gpurun_out/pytest_iter.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_iter.log 2> gpurun_out/bench_iter.err; echo "bench exit=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_iter.log').read().strip().splitlines()[-1])
    print("value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "fwd_tflops", round(d["fwd_tflops"],1), "clocks", d["clocks"])
    for k,v in d["kernels"].items(): print(k, {a:round(b,3) for a,b in v.items()})
    print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
except Exception as e:
    print("parse fail", e); print(open('gpurun_out/bench_iter.err').read()[-2000:])
PY
