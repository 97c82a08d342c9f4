"""Extracts the DRAM traffic of one kernel launch from an `ncu --set full` report and writes the small JSON that
bench.py reads for roofline.traffic (never a typed-in constant).

    python tools/ncu_traffic.py gpurun_out/prof.ncu-rep profiles/r02_traffic.json "kernel description" <algorithmic bytes>
"""
import csv
import io
import json
import subprocess
import sys

rep, out, desc, algo = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(hdr)}


def val(row, name):
    v = float(row[col[name]].replace(",", ""))
    u = units[col[name]].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


launches = []
for r in data:
    launches.append({"kernel": r[col["Kernel Name"]][:80], "dram_read_bytes": val(r, "dram__bytes_read.sum"),
                     "dram_write_bytes": val(r, "dram__bytes_write.sum"),
                     "duration_us": float(r[col["gpu__time_duration.sum"]].replace(",", "")) / (1e3 if units[col["gpu__time_duration.sum"]] in ("ns", "nsecond") else 1)})
best = max(launches, key=lambda d: d["dram_read_bytes"] + d["dram_write_bytes"])
best.update({"kernel": desc, "algorithmic_bytes": algo, "report": rep.split("/")[-1], "launches_in_report": len(launches)})
json.dump(best, open(out, "w"), indent=1)
print(json.dumps(best))
