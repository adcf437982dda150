#!/bin/bash
# One box with 8 GPUs: config 4 strong scaling (1/2/4/8), then bench.py weak scaling at 8 and 4.
mkdir -p gpurun_out
: > gpurun_out/scale_c4.log
for g in 1 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29600 + g)) \
      tools/multi_gpu_c4.py 2>> gpurun_out/scale_c4.err | grep '^{' | tee -a gpurun_out/scale_c4.log
done
for g in 8 4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29700 + g)) \
      bench.py --gpus $g --steps 10 --warmup 3 2>> gpurun_out/scale_bench.err | grep '^{' | tee gpurun_out/bench_n$g.log | cut -c1-400
done
# sequence-parallel (ring) attention at N = 65536 over all 8 GPUs, NCCL and peer-memory transports
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29811 \
    tools/multi_gpu_ring.py 1 8 65536 128 2>> gpurun_out/scale_ring.err | grep '^{' | tee gpurun_out/ring_n8.jsonl | cut -c1-300
