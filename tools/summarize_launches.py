"""Per-kernel shares of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/summarize_launches.py FILE [skip_first_n]"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(unit, 1e-6)
        rows.append((r["Kernel Name"], v * scale, int(r["ID"])))
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = rows[skip:]
    agg = defaultdict(lambda: [0, 0.0])
    for name, ms, _ in rows:
        name = re.sub(r"\(.*", "", name)
        agg[name][0] += 1
        agg[name][1] += ms
    tot = sum(v[1] for v in agg.values())
    for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:90]:90s} n={n:5d} time={ms:9.3f} ms share={100 * ms / tot:5.1f}%")
    print(f"TOTAL {tot:.3f} ms over {len(rows)} launches")


main()
