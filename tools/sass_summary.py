"""SASS mnemonic counts per kernel of a built library (python tools/sass_summary.py [e2enet_medical_b200/libe2enet_b200.so]):
the evidence that the contraction kernels are tcgen05 / TMEM / TMA code (profiles/r02_sass_summary.txt)."""
import os
import re
import subprocess
import sys
from collections import OrderedDict

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                           "e2enet_medical_b200", "libe2enet_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
COLS = OrderedDict([("UTCHMMA", r"\bUTCHMMA"), ("UTCBAR", r"\bUTCBAR"), ("LDTM", r"\bLDTM"), ("UTMALDG", r"\bUTMALDG"),
                    ("UTMAPF", r"\bUTMAPF"), ("UBLKCP", r"\bUBLKCP"), ("HMMA", r"\bHMMA"), ("SHFL", r"\bSHFL"),
                    ("BAR.SYNC", r"\bBAR\.SYNC"), ("ATOMG/RED", r"\b(ATOMG|RED)\b"), ("LDL/STL (spill)", r"\b(LDL|STL)\b")])
rows, name = OrderedDict(), None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "").replace("void ", ""))
        rows[name] = OrderedDict((c, 0) for c in COLS)
        continue
    if name and "/*" in line:
        for c, pat in COLS.items():
            if re.search(pat, line):
                rows[name][c] += 1
print("# SASS mnemonic counts per kernel of %s (cuobjdump -sass, sm_100a); UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit," % os.path.basename(lib))
print("# LDTM = tcgen05.ld, UTMALDG / UTMAPF = TMA tensor load / L2 prefetch, UBLKCP = bulk copy, HMMA = legacy mma.sync fallback family")
print("%-46s" % "kernel" + "".join("%10s" % c if len(c) <= 9 else "  " + c for c in COLS))
for n in sorted(rows):
    print("%-46s" % n[:46] + "".join("%10d" % v if len(c) <= 9 else "  %d" % v for c, v in rows[n].items()))
tot = OrderedDict((c, sum(r[c] for r in rows.values())) for c in COLS)
print("%-46s" % "TOTAL" + "".join("%10d" % v if len(c) <= 9 else "  %d" % v for c, v in tot.items()))
