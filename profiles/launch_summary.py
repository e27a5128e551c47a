"""Per-kernel totals of ONE spl_step inside an ncu launch list (the --csv --log-file output of
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,
smsp__inst_executed.sum --clock-control none`).  The step is the `which`-th one in the list,
from its k_advect_* launch to its gradient / check launch.
usage: python profiles/launch_summary.py profiles/r2_launches.csv [which=1]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr, data = rows[hi], rows[hi + 1:]
kn, mn, mv, idc, mu = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Metric Unit"))
launch = collections.OrderedDict()
for r in data:
    if len(r) <= mv:
        continue
    d = launch.setdefault(r[idc], {"name": r[kn]})
    v = float(r[mv].replace(",", ""))
    if r[mn] == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(r[mu], 1)      # -> microseconds
    d[r[mn]] = v
short = lambda n: re.sub(r"\(.*", "", n).replace("void ", "").replace("spl::", "").replace("<unnamed>::", "")
names = [short(d["name"]) for d in launch.values()]
vals = list(launch.values())
first = [i for i, n in enumerate(names)
         if n.startswith("k_advect") and not n.startswith("k_advect_rest")][which]
last = [i for i, n in enumerate(names) if i > first and n.startswith(("k_grad_check4", "k_check"))][0]
tot, cnt, byt, ins = (collections.Counter() for _ in range(4))
for i in range(first, last + 1):
    d, n = vals[i], names[i]
    tot[n] += d["gpu__time_duration.sum"]
    cnt[n] += 1
    byt[n] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
    ins[n] += d.get("smsp__inst_executed.sum", 0)
T = sum(tot.values())
print(f"# one spl_step: launches {first}..{last} of {path} (times under ncu are cold-cache and "
      f"serialised: read the shares, not the absolutes)")
print("kernel | launches | total us | share | us/launch | DRAM MB/launch | warp instr/launch")
for n, t in tot.most_common():
    print(f"{n} | {cnt[n]} | {t:.1f} | {t / T:.3f} | {t / cnt[n]:.1f} | {byt[n] / cnt[n] / 1e6:.1f} | {ins[n] / cnt[n]:.3g}")
