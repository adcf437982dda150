#!/bin/bash
# full validation of the cleaned-up library + preprocess grid sweep + the competitor sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
for c in 2 4 8 16 32 64; do FA_PRE_CTAS_PER_SM=$c timeout 100 python tools/kernel_times.py 2>&1 | grep "N8192" | sed "s/^/pre_ctas_per_sm=$c /"; done | tee gpurun_out/pre_sweep.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), {k: round(v["ms"], 4) for k, v in d["kernels"].items()})
print("e2e", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in d["e2e"].items() if k != "api"})
print("c4", d["c4_strong"]); print("cpu", d["cpu_baseline"]); print("clocks", d["clocks"])
PY
timeout 1500 python tools/bench_sweep.py > gpurun_out/sweep.log 2>&1; echo "sweep exit=$?"; grep -v Warn gpurun_out/sweep.log | tail -70
