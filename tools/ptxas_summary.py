"""summarises a ptxas -v log: registers / spills per kernel (python tools/ptxas_summary.py e2enet_medical_b200/_build/conv_tc.o.log)"""
import re
import subprocess
import sys

name, rows = None, []
for line in open(sys.argv[1]):
    m = re.search(r"Compiling entry function '([^']+)'", line)
    if m:
        name = m.group(1)
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m:
        spill = (m.group(2), m.group(3))
    m = re.search(r"Used (\d+) registers", line)
    if m and name:
        d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        d = re.sub(r"\(.*", "", d.replace("(anonymous namespace)::", "").replace("void ", ""))
        rows.append((d, int(m.group(1)), spill))
        name = None
for d, r, sp in rows:
    print("%-44s regs %4d  spill stores %4s B  loads %4s B" % (d, r, sp[0], sp[1]))
