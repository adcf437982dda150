#!/bin/bash
mkdir -p gpurun_out
export FA_B200_LIB=build/var/libfa_w16wd.so
FA_FWD_W16=0 FA_PROBE_SAVE=/tmp/fa_w8 timeout 300 python tools/fwd_pair_probe.py > gpurun_out/w16_0.log 2>&1; echo "w16=0 exit=$?"
FA_FWD_W16=1 FA_PROBE_COMPARE=/tmp/fa_w8 timeout 300 python tools/fwd_pair_probe.py > gpurun_out/w16_1.log 2>&1; echo "w16=1 exit=$?"
grep -v Warn gpurun_out/w16_1.log | tail -100
echo ---- eight warps
grep -v Warn gpurun_out/w16_0.log | grep -E "N8192|N32768|B8|B4"
