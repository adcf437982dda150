#!/bin/bash
mkdir -p gpurun_out
for v in f2noeg f2unord f2ord; do
  FA_B200_LIB=$PWD/build/var/libfa_$v.so timeout 200 python tools/fused_probe.py 3 2>&1 | grep "^lib"
done | tee gpurun_out/fused2.txt
FA_B200_LIB=$PWD/build/var/libfa_f2ord.so timeout 600 python -m pytest tests/test_gpu_parity.py -q -x --timeout 300 -m gpu \
  -k "fused or variants_bit_identical or graph_capture" > gpurun_out/fused2_pytest.log 2>&1
echo "pytest exit=$?"; tail -2 gpurun_out/fused2_pytest.log
