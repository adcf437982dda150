#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -2 gpurun_out/pytest_gpu.log
for i in 1 2; do timeout 120 python tools/kernel_times.py 2>&1 | grep "^lib"; done | tee gpurun_out/split_prod.txt
