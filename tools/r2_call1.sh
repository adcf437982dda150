#!/bin/bash
# Round 2, GPU call 1: full parity suite on the hygiene changes, baseline bench (watchdog off vs on), measured parity
# errors, and the single-pass backward ablations (turn warps, half egress).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.log 2>&1
timeout 900 python -m pytest tests -q -m gpu -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_wdoff.log 2> gpurun_out/bench_wdoff.err; echo "bench exit=$?"
FA_B200_LIB=build/var/libfa_wd.so timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_wdon.log 2> gpurun_out/bench_wdon.err; echo "bench(wd) exit=$?"
python - <<'PY'
import json
for tag in ("wdoff", "wdon"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{tag}.log").read().strip().splitlines()[-1])
        print(tag, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), {k: round(v["ms"], 4) for k, v in d["kernels"].items()},
              "e2e", round(d["e2e"]["value"], 1), d["clocks"])
    except Exception as e:
        print(tag, "parse fail", e)
PY
for v in turn r1order unord noegress halfunord halford halfunord_nopace; do
  FA_B200_LIB=build/var/libfa_$v.so timeout 200 python tools/fused_probe.py 2 2>&1 | grep -v Warning
done | tee gpurun_out/fused_variants.log
timeout 600 python tools/parity_errors.py gpurun_out/parity_errors.json > gpurun_out/parity_errors.log 2>&1; echo "parity_errors exit=$?"; tail -20 gpurun_out/parity_errors.log
timeout 200 python tools/host_overhead_probe.py > gpurun_out/host_overhead.log 2>&1; head -4 gpurun_out/host_overhead.log
timeout 200 python tools/small_probe.py > gpurun_out/small.log 2>&1; cat gpurun_out/small.log
