#!/bin/bash
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1500 python tools/bench_sweep.py > gpurun_out/sweep.log 2>&1; echo "sweep exit=$? after $(( $(date +%s) - t0 )) s"; tail -3 gpurun_out/sweep.log
timeout 200 python tools/parity_errors.py > gpurun_out/parity_errors.log 2>&1; echo "parity exit=$?"; tail -2 gpurun_out/parity_errors.log
