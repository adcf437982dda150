#!/bin/bash
# ncu evidence for profiles/: launch list of a short bench run + one full-set capture of each of our kernels
# (the second step of tools/prof_step.py; with the optional single-pass backward appended: 7 kernels per step).
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit=$?"
python tools/prof_step.py fused > gpurun_out/plain_step.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fa_ -s 7 -c 7 -f -o gpurun_out/prof_step \
    python tools/prof_step.py fused > gpurun_out/ncu_full.log 2>&1
echo "full capture exit=$?"
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/
