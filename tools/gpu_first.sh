#!/bin/bash
# bring-up: run test groups separately so one hang / fault does not hide the rest
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python -m pytest tests/test_gpu_parity.py -q --timeout 300 "$@" > gpurun_out/t_$name.log 2>&1; echo "exit=$?" >> gpurun_out/t_$name.log; echo "== $name"; tail -n 25 gpurun_out/t_$name.log | cut -c1-300; }
run fp32 -k "fp32"
run pre -k "preprocess or errors"
run fwd16 -k "test_16bit_parity and 128-False-64-dtype0"
run p16 -k "test_16bit_parity"
run rest -k "not fp32 and not test_16bit_parity and not preprocess and not errors and not config"
run cfg -k "config"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit=$?" >> gpurun_out/bench.err
cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
