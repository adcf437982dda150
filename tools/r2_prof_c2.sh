#!/bin/bash
mkdir -p gpurun_out
python tools/prof_c2.py > gpurun_out/prof_c2_plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/prof_c2_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fa_fwd_kernel|fa_bwd_d" -s 3 -c 3 -f -o gpurun_out/prof_c2 \
    python tools/prof_c2.py > gpurun_out/ncu_c2.log 2>&1
echo "ncu exit=$?"; tail -3 gpurun_out/ncu_c2.log; ls -la gpurun_out/prof_c2*
