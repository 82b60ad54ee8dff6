"""Per-kernel SASS evidence of the built library: counts of the mnemonics that prove the sm_100a code paths
(DMMA = FP64 tensor pipe via mma.sync f64, UTMALDG = TMA loads, FENCE.VIEW.ASYNC = cross-proxy fence of the stage
release, SYNCS = mbarrier operations, REDUX/MATCH = warp vote primitives).

    python tools/sass_summary.py > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "scregclust_b200", "lib", "libscregclust_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
want = ["DMMA", "UTMALDG", "FENCE.VIEW.ASYNC", "SYNCS", "LDS", "LDG", "STG", "DFMA", "DADD", "DMUL", "SHFL", "REDUX", "MATCH", "ATOMS", "BAR"]
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("scrc::", "").replace("void ", "")
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        for w in want:
            if op == w or op.startswith(w + "."):
                counts[cur][w] += 1
                total[w] += 1
        counts[cur]["_all"] += 1
print(f"cuobjdump -sass {os.path.relpath(lib, ROOT)}  (sm_100a; built with -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false)")
print("totals: " + ", ".join(f"{w} {total[w]}" for w in want))
print(f"{'kernel':78s} {'instr':>7s} " + " ".join(f"{w.split('.')[0][:7]:>7s}" for w in want))
for k, c in counts.items():
    if c["_all"] == 0:
        continue
    print(f"{k[:78]:78s} {c['_all']:7d} " + " ".join(f"{c[w]:7d}" for w in want))
