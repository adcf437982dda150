#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x --timeout 600 -m gpu -k "backward or bwd or grad or parity or dropout or mask or seqlens or randomized or reference" > gpurun_out/dqfull_pytest.log 2>&1
echo "pytest exit=$?"; tail -3 gpurun_out/dqfull_pytest.log
for i in 1 2 3; do
  for v in dqhalf ""; do
    if [ -n "$v" ]; then export FA_B200_LIB=$PWD/build/var/libfa_$v.so; else unset FA_B200_LIB; fi
    timeout 120 python tools/kernel_times.py 2>&1 | grep "^lib" | grep D128
  done
done | tee gpurun_out/dqfull.txt
