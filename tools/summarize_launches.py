"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean ns, share.

    python tools/summarize_launches.py gpurun_out/launches.csv [first_launch last_launch]
"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        i = int(r["ID"])
        if lo <= i <= hi:
            rows.append((i, r["Kernel Name"], r["Grid Size"], r["Block Size"], float(r["Metric Value"])))
    agg = OrderedDict()
    for i, name, grid, block, ns in rows:
        short = re.sub(r"\(.*", "", name)
        short = re.sub(r"^void ", "", short)
        a = agg.setdefault(short, {"n": 0, "ns": 0.0, "grid": grid, "block": block})
        a["n"] += 1
        a["ns"] += ns
    total = sum(a["ns"] for a in agg.values())
    print(f"launches {lo}..{min(hi, rows[-1][0])}: {len(rows)} kernels, {total / 1e3:.1f} us of GPU time (cold-cache, serialised)")
    print("| kernel | launches | mean us | share | grid | block |")
    print("|---|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        print(f"| `{k[:90]}` | {a['n']} | {a['ns'] / a['n'] / 1e3:.1f} | {100 * a['ns'] / total:.1f}% | {a['grid']} | {a['block']} |")


if __name__ == "__main__":
    main()
