#!/bin/bash
mkdir -p gpurun_out
export FA_B200_LIB=build/var/libfa_pairwd.so
FA_FWD_PAIR=1 timeout 300 python tools/fwd_pair_probe.py > gpurun_out/pair1.log 2>&1; echo "pair=1 exit=$?"
FA_FWD_PAIR=0 timeout 300 python tools/fwd_pair_probe.py > gpurun_out/pair0.log 2>&1; echo "pair=0 exit=$?"
grep -v Warn gpurun_out/pair1.log | tail -60
python - <<'PY'
a = [l.split() for l in open("gpurun_out/pair0.log") if l.startswith("pair=")]
b = [l.split() for l in open("gpurun_out/pair1.log") if l.startswith("pair=")]
bad = 0
for x, y in zip(a, b):
    same = x[8] == y[8]
    bad += not same
    print(" ".join(x[1:8]), "single", x[9], "ms | pair", y[9], "ms", "bits equal" if same else "BITS DIFFER")
print("shapes compared", min(len(a), len(b)), "of", len(a), len(b), "mismatches", bad)
PY
