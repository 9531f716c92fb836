"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel."""
import csv
import collections
import sys


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        rows.append((r["Kernel Name"], v * scale))
    tot = sum(t for _, t in rows) or 1.0
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, t in rows:
        name = k.split("(")[0][:70]
        agg[name][0] += 1
        agg[name][1] += t
    print(f"{len(rows)} launches, {tot/1e3:.3f} ms total (cold-cache, serialised: compare shares)")
    print(f"{'kernel':72s} {'launches':>8s} {'us total':>12s} {'share':>7s} {'us/launch':>10s}")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:72s} {n:8d} {t:12.1f} {100*t/tot:6.1f}% {t/n:10.2f}")


if __name__ == "__main__":
    main(sys.argv[1])
