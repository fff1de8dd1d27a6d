"""Counts, per kernel of libsregex_cuda.so, the SASS instructions that show how it moves data:
UTMALDG (TMA tile loads), SYNCS (mbarrier), REDUX (warp reduction), LDS / LDG / STS / STG, FENCE.
  python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "sregex_b200", "libsregex_cuda.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
keys = ["UTMALDG", "SYNCS", "REDUX", "FENCE", "LDS", "LDG", "STS", "STG", "PRMT", "HMMA", "UTC"]
kern, counts, arch = None, collections.OrderedDict(), set()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        full = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        km = re.search(r"(k_\w+(?:<[^>]*>)?)", full)
        kern = km.group(1) if km else full[:70]
        counts.setdefault(kern, collections.Counter())
        continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch.add(m.group(1))
    if kern:
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            counts[kern]["total"] += 1
            for k in keys:
                if op.startswith(k):
                    counts[kern][k] += 1
print(f"# {os.path.relpath(so, ROOT)}: cubins for {sorted(arch)}; instruction counts per kernel (cuobjdump -sass)")
print(f"{'kernel':70s} " + " ".join(f"{k:>8s}" for k in ["total"] + keys))
for k, c in counts.items():
    print(f"{k[:70]:70s} " + " ".join(f"{c.get(x, 0):8d}" for x in ["total"] + keys))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print(f"{'ALL':70s} " + " ".join(f"{tot.get(x, 0):8d}" for x in ["total"] + keys))
