"""SASS census of the shipped library: per kernel, how many TMA loads (UTMALDG), mbarrier
operations (SYNCS), cp.async (LDGSTS), shared / global memory instructions, packed f32x2 and
fp64 arithmetic.  usage: python profiles/sass_census.py > profiles/r2_sass_census.txt
(cuobjdump only; no GPU needed)."""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "splashy_b200" / "lib" / "libsplashy_cuda.so"
txt = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
funcs = re.split(r"\n\s*Function : ", txt)
names = [f.split("\n", 1)[0].strip() for f in funcs[1:]]
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
rows = []
for f, d in zip(funcs[1:], dem):
    m = re.match(r"(?:void )?(?:spl::)?([\w:]+(?:<[^(]*>)?)\(", d)
    short = m.group(1) if m else d[:60]
    c = collections.Counter(re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", f, re.M))
    rows.append((short, sum(c.values()), c["UTMALDG"], c["SYNCS"], c["LDGSTS"], c["LDS"], c["STS"],
                 c["LDG"], c["STG"], c["BAR"], c["SHFL"], c["FFMA2"] + c["FADD2"] + c["FMUL2"],
                 c["DFMA"] + c["DADD"] + c["DMUL"]))
hdr = ("kernel", "instr", "UTMALDG", "SYNCS", "LDGSTS", "LDS", "STS", "LDG", "STG", "BAR", "SHFL",
       "f32x2", "fp64")
print(f"# {LIB.name}: architectures {arch}; {len(rows)} kernels")
print("# UTMALDG = TMA tensor load (cp.async.bulk.tensor), SYNCS = mbarrier operations, LDGSTS = cp.async,")
print("# f32x2 = FFMA2 + FADD2 + FMUL2 (packed pairs), fp64 = DFMA + DADD + DMUL")
print(" | ".join(hdr))
for r in sorted(rows, key=lambda r: (-r[2], -r[4], r[0])):
    print(" | ".join(str(x) for x in r))
