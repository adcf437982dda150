#!/bin/bash
mkdir -p gpurun_out
timeout 20 ./build/softmax_tlp_probe > gpurun_out/softmax_tlp_probe.txt 2>&1; echo "exit=$?"; tail -3 gpurun_out/softmax_tlp_probe.txt
timeout 20 ./build/softmax_tlp_probe bg > gpurun_out/softmax_tlp_probe_bg.txt 2>&1; echo "exit=$?"; cat gpurun_out/softmax_tlp_probe_bg.txt
nvidia-smi --query-gpu=name,memory.used --format=csv
