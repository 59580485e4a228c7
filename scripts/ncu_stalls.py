"""Summarise `ncu --page source --csv` output: top stall reasons and hottest SASS lines.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K --launch-count 1 > src.csv
       python scripts/ncu_stalls.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] != "Address"]
S = idx["# Samples"]
tot = sum(int(r[S]) for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[idx[s]]) for r in data) for s in stalls}
print("total samples", tot)
print("stall totals:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for r in sorted(data, key=lambda r: -int(r[S]))[:top_n]:
    best = max(stalls, key=lambda s: int(r[idx[s]]))
    print(f"{int(r[S]):7d} {100.0 * int(r[S]) / tot:5.1f}%  {r[idx['Source']].strip()[:64]:64s} {best[6:]:12s} exec={r[idx['Instructions Executed']]}")
