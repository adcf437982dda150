#!/bin/bash
mkdir -p gpurun_out
FA_B200_LIB=$PWD/build/var/libfa_kvtmem.so timeout 600 python -m pytest tests/test_gpu_parity.py -q -x --timeout 300 -m gpu \
  -k "two_kernel_backward or bit_identical or reference_correctness or randomized" > gpurun_out/kvtmem_pytest.log 2>&1
echo "pytest exit=$?"; tail -2 gpurun_out/kvtmem_pytest.log
for i in 1 2; do
  for v in slots2 kvtmem ""; do
    if [ -n "$v" ]; then export FA_B200_LIB=$PWD/build/var/libfa_$v.so; else unset FA_B200_LIB; fi
    timeout 120 python tools/kernel_times.py 2>&1 | grep "^lib" | grep D64
  done
done | tee gpurun_out/kvtmem.txt
