"""Per-kernel shares of ONE training step from an ncu launch list of tools/bench_train.py (NO_GRAPH=1): the launches
between two consecutive optimiser kernels.  python tools/step_summary.py LAUNCHES.csv [step_index]"""
import collections
import csv
import re
import sys


def main():
    with open(sys.argv[1]) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        rows.append((re.sub(r"\(.*", "", r["Kernel Name"]), float(r["Metric Value"].replace(",", "")) / 1e3))
    idx = [i for i, r in enumerate(rows) if "adamw" in r[0]]
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    step = rows[idx[k] + 1:idx[k + 1] + 1] if len(idx) > k + 1 else rows   # a list already trimmed to one step
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, t in step:
        agg[n][0] += 1
        agg[n][1] += t
    tot = sum(v[1] for v in agg.values())
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n[:86]:86s} n={c:4d} time={t / 1e3:7.3f} ms share={100 * t / tot:5.1f}%")
    print(f"TOTAL {tot / 1e3:.3f} ms over {len(step)} launches (serialised, cold-cache ncu durations: compare shares)")


main()
