"""Per-kernel SASS opcode histogram of the built library (cuobjdump -sass), so that the Blackwell/TMA/cp.async/FP64 claims of DESIGN.md
can be checked without rebuilding:  python profiles/sass_summary.py > profiles/sass_summary.txt
Columns: total instructions, DFMA (+ DADD/DMUL), IDP.4A (dp4a byte sums), UBLKCP (cp.async.bulk = TMA bulk copy, both directions),
LDGSTS (cp.async), SYNCS (mbarrier), LDS/STS, SHFL, BAR, and whether tensor-core / legacy-MMA opcodes appear (they must not)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multi-rtl-sdr-calibration_b200", "csrc", "libgsmcal.so")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
arch = set(re.findall(r"arch = (sm_\w+)", out))
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "")
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
    if m and kern:
        op = m.group(1)
        hist[kern]["total"] += 1
        base = op.split(".")[0]
        hist[kern][base] += 1
        if op.startswith("IDP.4A"):
            hist[kern]["IDP.4A"] += 1
cols = ["total", "DFMA", "DADD", "DMUL", "IDP.4A", "UBLKCP", "LDGSTS", "SYNCS", "LDS", "STS", "SHFL", "BAR"]
print("# cuobjdump -sass of libgsmcal.so, arch %s; opcode counts per kernel (static code)" % ",".join(sorted(arch)))
print("%-44s " % "kernel" + " ".join("%7s" % c for c in cols) + "  tensor/MMA opcodes")
for k, h in hist.items():
    mma = sorted(o for o in h if re.match(r"(HMMA|IMMA|DMMA|HGMMA|QGMMA|IGMMA|UTC.*MMA|LDTM|STTM)", o))
    print("%-44s " % k[:44] + " ".join("%7d" % h.get(c, 0) for c in cols) + "  " + (",".join(mma) if mma else "none"))
tot = collections.Counter()
for h in hist.values():
    tot.update(h)
print("%-44s " % "ALL KERNELS" + " ".join("%7d" % tot.get(c, 0) for c in cols))
