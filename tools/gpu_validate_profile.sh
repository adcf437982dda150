#!/bin/bash
# final-ish validation: full GPU suite, smoke, bench (with CPU leg) + reference arm, ncu launch list + full capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit=$?"; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "bench ref exit=$?"; cat gpurun_out/bench_ref.log | cut -c1-300
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), {k: round(v["ms"], 4) for k, v in d["kernels"].items()})
print("roofline", d["roofline"]); print("e2e", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in d["e2e"].items() if k not in ("api", "pcie_bare_note")})
print("clocks", d["clocks"])
PY
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit=$?"
python tools/prof_step.py > gpurun_out/plain_step.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fa_ -s 4 -c 4 -f -o gpurun_out/prof_step \
    python tools/prof_step.py > gpurun_out/ncu_full.log 2>&1
echo "full capture exit=$?"; tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/prof_step.ncu-rep
