"""SASS digest of the in-tree CUDA library: per kernel, how many tcgen05 MMA (UTCHMMA, .2CTA), TMA load / store
(UTMALDG / UTMASTG, .MULTICAST), TMEM load (LDTM) and tensor-memory barrier (UTCBAR) instructions it holds.
Usage: python scripts/sass_digest.py [path/to/lib.so] > profiles/rN_sass_digest.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "cianna_b200", "libcianna_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True, check=True).stdout
pats = ["UTCHMMA.2CTA", "UTCHMMA", "UTMALDG", "MULTICAST", "UTMASTG", "LDTM", "UTCBAR", "SYNCS", "ATOMS.CAST", "RED.E.ADD.F64", "RED.E.ADD.F32"]
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None or "/*" not in line:
        continue
    per[cur]["instr"] += 1
    for p in pats:
        if p in line:
            per[cur][p] += 1
demangled = subprocess.run(["c++filt"], input="\n".join(per), stdout=subprocess.PIPE, text=True).stdout.splitlines()
tot = collections.Counter()
print("# %s  (%d kernels)" % (os.path.relpath(so, ROOT), len(per)))
print("# columns: instructions | UTCHMMA (of which .2CTA) | UTMALDG (of which .MULTICAST) | UTMASTG | LDTM | UTCBAR")
for (name, c), dn in zip(per.items(), demangled):
    tot.update(c)
    if c["UTCHMMA"] or c["UTMALDG"] or c["UTMASTG"] or c["LDTM"]:
        short = re.sub(r"\(.*", "", dn.replace("void ", "").replace("cb200::", ""))
        print("%-72s %6d | %4d (%3d) | %4d (%3d) | %3d | %3d | %3d" % (short[:72], c["instr"], c["UTCHMMA"], c["UTCHMMA.2CTA"], c["UTMALDG"],
              c["MULTICAST"], c["UTMASTG"], c["LDTM"], c["UTCBAR"]))
print("# total: %d instructions, UTCHMMA %d (.2CTA %d), UTMALDG %d (.MULTICAST %d), UTMASTG %d, LDTM %d, UTCBAR %d; kernels without tensor / TMA instructions: %d"
      % (tot["instr"], tot["UTCHMMA"], tot["UTCHMMA.2CTA"], tot["UTMALDG"], tot["MULTICAST"], tot["UTMASTG"], tot["LDTM"], tot["UTCBAR"],
         sum(1 for c in per.values() if not (c["UTCHMMA"] or c["UTMALDG"] or c["UTMASTG"] or c["LDTM"]))))
