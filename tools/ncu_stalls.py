"""Summarise the per-instruction stall samples of one kernel from `ncu --page source --csv` output.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv ; python tools/ncu_stalls.py src.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr, data = rows[hi], rows[hi + 1:]
ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot, agg, recs = 0, {s: 0 for s in stalls}, []
for r in data:
    if len(r) < len(hdr) or not r[ci["# Samples"]].isdigit():
        continue
    n = int(r[ci["# Samples"]])
    tot += n
    st = {s: int(r[ci[s]] or 0) for s in stalls}
    for s in stalls:
        agg[s] += st[s]
    recs.append((n, r[ci["Address"]][-5:], r[ci["Source"]].strip(), {k: v for k, v in st.items() if v}))
print("total samples", tot)
print(sorted(agg.items(), key=lambda kv: -kv[1])[:10])
for n, a, src, st in sorted(recs, key=lambda x: -x[0])[:top]:
    print(n, a, src[:72], dict(sorted(st.items(), key=lambda kv: -kv[1])[:3]))
