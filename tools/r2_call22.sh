#!/bin/bash
# 8 GPUs: bench.py --gpus 8 (weak scaling, c4_strong, e2e, selftest) and the ring tool at N = 65536
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.log 2> gpurun_out/bench_n8.err; echo "bench n8 exit=$?"; tail -3 gpurun_out/bench_n8.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_n8.log") if l.startswith("{")][-1])
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3))
print("e2e", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in d["e2e"].items() if k not in ("api", "pcie_bare_note")})
print("c4", d["c4_strong"]); print("selftest", json.dumps(d["multi_gpu_selftest"]))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 tools/multi_gpu_ring.py 1 8 65536 128 2>&1 | grep "^{" | tee gpurun_out/ring_8gpu.log
