#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./build/umma2_probe > gpurun_out/umma2_probe.log 2>&1; echo "probe exit=$?"; cat gpurun_out/umma2_probe.log
for i in 1 2; do
  timeout 100 python tools/kernel_times.py 2>&1 | grep lib
  FA_B200_LIB=build/var/libfa_wd.so timeout 100 python tools/kernel_times.py 2>&1 | grep lib
done | tee gpurun_out/kernel_times.log
timeout 600 python -m pytest tests -q -m gpu -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
