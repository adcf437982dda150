#!/bin/bash
# Single-pass backward with the TMA-staged dQ egress: parity of the ordered build (watchdog on), then timings of the variants.
mkdir -p gpurun_out
FA_B200_LIB=$PWD/build/var/libfa_tmaord.so timeout 600 python -m pytest tests/test_gpu_parity.py -q -x --timeout 300 -m gpu \
  -k "fused or variants_bit_identical or graph_capture" > gpurun_out/fused_tma_pytest.log 2>&1
echo "pytest exit=$?"; tail -5 gpurun_out/fused_tma_pytest.log
for v in tmanoeg tmaunord tmaordp oldord; do
  FA_B200_LIB=$PWD/build/var/libfa_$v.so timeout 200 python tools/fused_probe.py 3 2>&1 | grep "^lib"
done | tee gpurun_out/fused_tma.txt
