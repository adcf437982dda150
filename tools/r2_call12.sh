#!/bin/bash
mkdir -p gpurun_out
FA_B200_LIB=build/var/libfa_dq1wd.so timeout 300 python tools/bwd_ab_probe.py 2>&1 | grep "^lib" > gpurun_out/dq1.log; echo "one issuer exit=$?"
FA_B200_LIB=build/var/libfa_dq2wd.so timeout 300 python tools/bwd_ab_probe.py 2>&1 | grep "^lib" > gpurun_out/dq2.log; echo "two issuers exit=$?"
python - <<'PY'
a = [l.split() for l in open("gpurun_out/dq1.log")]
b = [l.split() for l in open("gpurun_out/dq2.log")]
bad = sum(x[8] != y[8] for x, y in zip(a, b))
print("shapes", len(a), len(b), "checksum mismatches", bad)
for x, y in zip(a, b):
    if float(x[12]) > 0:
        print(" ".join(x[2:8]), "one issuer: dkdv", x[10], "dq", x[12], "| two issuers: dkdv", y[10], "dq", y[12], "" if x[8] == y[8] else "BITS DIFFER")
PY
