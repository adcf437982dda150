#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./build/umma2_probe > gpurun_out/umma2_probe.log 2>&1; echo "probe exit=$?"; grep -E "grid 148|pair case" gpurun_out/umma2_probe.log
export FA_B200_LIB=build/var/libfa_pairtrace.so
FA_TRACE_CAUSAL=0 timeout 200 python tools/trace_fwd2.py 2>&1 | grep -v Warn | tee gpurun_out/trace_fwd2.log
