#!/bin/bash
mkdir -p gpurun_out
FA_B200_LIB=build/var/libfa_dec0wd.so FA_PROBE_SAVE=/tmp/fa_d0 timeout 300 python tools/fwd_pair_probe.py > gpurun_out/dec0.log 2>&1; echo "off exit=$?"
FA_B200_LIB=build/var/libfa_dec1wd.so FA_PROBE_COMPARE=/tmp/fa_d0 timeout 300 python tools/fwd_pair_probe.py > gpurun_out/dec1.log 2>&1; echo "on exit=$?"
echo "bit-equal shapes:" $(grep -c "bits equal" gpurun_out/dec1.log) "of" $(grep -c "^pair" gpurun_out/dec1.log); grep "^pair" gpurun_out/dec1.log | grep -v "bits equal" | head
echo ---- timings off vs on
paste <(grep -E "N8192|N32768|B8 |B4 " gpurun_out/dec0.log | awk '{print $2,$3,$4,$5,$6,$7,$9,$11}') <(grep -E "N8192|N32768|B8 |B4 " gpurun_out/dec1.log | awk '{print $9,$11}')
