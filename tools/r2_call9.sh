#!/bin/bash
# 2 GPUs: the two-GPU parity tests, the ring tool (side-stream prefetch), and bench.py --gpus 2 (selftest, c4_strong, e2e)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x --timeout 600 -k "two_gpus or single_rank or ring" > gpurun_out/pytest_2gpu.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_2gpu.log; tail -4 gpurun_out/pytest_2gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/multi_gpu_ring.py 1 8 32768 128 2>&1 | grep "^{" | tee gpurun_out/ring_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "bench n2 exit=$?"; tail -3 gpurun_out/bench_n2.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_n2.log") if l.startswith("{")][-1])
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3))
print("e2e", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in d["e2e"].items() if k not in ("api", "pcie_bare_note")})
print("c4", d["c4_strong"]); print("selftest", json.dumps(d["multi_gpu_selftest"], indent=1))
PY
