"""Summarise `ncu --page source --csv` output: stall mix and hottest SASS lines.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K > src.csv; python ncu_src.py src.csv"""
import csv
import sys


def num(x):
    try:
        return int(float(x))
    except Exception:
        return 0


rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if r and r[0] == "Address")
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows if len(r) == len(hdr) and r[0] != "Address"]
seen, inst = set(), []
for r in data:                       # first kernel instance only
    if r[ix["Address"]] in seen:
        break
    seen.add(r[ix["Address"]])
    inst.append(r)
tot = sum(num(r[ix["# Samples"]]) for r in inst)
print("instructions", len(inst), "samples", tot)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(num(r[ix[s]]) for r in inst) for s in stalls}
print("stall mix:", [(k, round(100.0 * v / max(tot, 1), 1)) for k, v in
                     sorted(agg.items(), key=lambda x: -x[1])[:8]])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
for r in sorted(inst, key=lambda r: -num(r[ix["# Samples"]]))[:n]:
    why = {s[6:]: num(r[ix[s]]) for s in stalls if num(r[ix[s]]) > 0.15 * max(1, num(r[ix["# Samples"]]))}
    print(f'{num(r[ix["# Samples"]]):7d}  {r[ix["Source"]][:86]:86s} {why}')
