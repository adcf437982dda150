#!/bin/bash
mkdir -p gpurun_out
FA_B200_LIB=build/var/libfa_nosplitwd.so FA_PROBE_SAVE=/tmp/fa_ns timeout 300 python tools/fwd_pair_probe.py > gpurun_out/split0.log 2>&1; echo "nosplit exit=$?"
FA_B200_LIB=build/var/libfa_splitwd.so FA_PROBE_COMPARE=/tmp/fa_ns timeout 300 python tools/fwd_pair_probe.py > gpurun_out/split1.log 2>&1; echo "split exit=$?"
grep -c "bits equal" gpurun_out/split1.log; grep -v Warn gpurun_out/split1.log | grep -v "bits equal" | head -20
echo ---- timings split vs nosplit
paste <(grep -E "N8192|N32768|B8 |B4 " gpurun_out/split0.log | awk '{print $2,$3,$4,$5,$6,$7,$9,$11}') <(grep -E "N8192|N32768|B8 |B4 " gpurun_out/split1.log | awk '{print $9,$11}')
