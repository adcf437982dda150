#!/bin/bash
mkdir -p gpurun_out
for v in default fp00 fp92 fpaa bp88 bp92 default; do
  if [ $v = default ]; then unset FA_B200_LIB; else export FA_B200_LIB=build/var/libfa_$v.so; fi
  timeout 100 python tools/kernel_times.py 2>&1 | grep "^lib"
done | tee gpurun_out/poly_sweep.log
