#!/bin/bash
# Dev-container only: place an UNMODIFIED copy of the reference's Python sources under baseline/_ref/ (git-ignored,
# travels to the GPU box with gpurun) so tools/bench_sweep.py can time the reference Triton kernel on the B200.
set -e
cd "$(dirname "$0")/.."
mkdir -p baseline/_ref
rm -rf baseline/_ref/src
cp -r /root/reference/src baseline/_ref/src
ls baseline/_ref/src
