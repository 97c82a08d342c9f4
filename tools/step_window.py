"""Aggregates ONE training step out of an `ncu --metrics gpu__time_duration.sum --csv` launch list: the launches
between two consecutive pack_weights_multi_kernel launches (that kernel runs exactly once per step)."""
import csv, gzip, re, sys
from collections import defaultdict

def load(path):
    op = gzip.open if path.endswith('.gz') else open
    with op(path, 'rt') as f:
        lines = [l for l in f if l.startswith('"')]
    rows = []
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        n = re.sub(r"\(.*", "", r["Kernel Name"])
        n = n.replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        rows.append((n, v / 1000 if u in ("ns", "nsecond") else v))
    return rows

def one_step(rows):
    idx = [i for i, (n, _) in enumerate(rows) if n.startswith("pack_weights_multi")]
    if len(idx) >= 2:
        return rows[idx[0]:idx[1]]
    return rows

if __name__ == "__main__":
    rows = one_step(load(sys.argv[1]))
    tot = defaultdict(lambda: [0, 0.0])
    for n, us in rows:
        tot[n][0] += 1
        tot[n][1] += us
    total = sum(v[1] for v in tot.values())
    print(f"one step: {total / 1000:.2f} ms over {len(rows)} launches")
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
    for n, (c, us) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{us / 1000:9.3f} ms {100 * us / total:5.1f}% {c:5d}x  {n[:100]}")
