#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x --timeout 600 -m gpu -k "cpp_autograd or autograd or reference or gradcheck or errors or graph or streams" > gpurun_out/cpp_pytest.log 2>&1
echo "pytest exit=$?"; tail -3 gpurun_out/cpp_pytest.log
timeout 300 python tools/small_sweep_probe.py 2>&1 | tee gpurun_out/small_cpp.txt
FA_PROBE_PY=1 timeout 300 python tools/small_sweep_probe.py 2>&1 | tee gpurun_out/small_py.txt
