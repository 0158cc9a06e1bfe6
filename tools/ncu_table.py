"""Per-kernel table from an `ncu --page raw --csv` export:  python tools/ncu_table.py gpurun_out/r2_all_raw.csv [--md]"""
import csv, re, sys
path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr = rows[0]
col = {k: i for i, k in enumerate(hdr)}
def num(r, k):
    try: return float(r[col[k]].replace(",", ""))
    except Exception: return float("nan")
want = [("us", "gpu__time_duration.sum", 1e-3), ("dram rd MB", "dram__bytes_read.sum", 1e-6), ("dram wr MB", "dram__bytes_write.sum", 1e-6),
        ("dram %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1), ("sm %", "sm__throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1), ("occ %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
        ("regs", "launch__registers_per_thread", 1), ("grid", "launch__grid_size", 1), ("block", "launch__block_size", 1),
        ("smem KB", "launch__shared_mem_per_block_dynamic", 1e-3)]
units = rows[1]
out = []
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"msvit::", "", name)
    vals = []
    for label, key, sc in want:
        v = num(r, key) if key in col else float("nan")
        u = units[col[key]] if key in col else ""
        if label == "us" and u == "us": sc = 1
        if label == "us" and u == "ms": sc = 1e3
        if label == "us" and u == "ns": sc = 1e-3
        if label.startswith("dram") and label.endswith("MB"):
            sc = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, sc)
        if label == "smem KB":
            sc = {"byte/block": 1e-3, "Kbyte/block": 1}.get(u, sc)
        vals.append(v * sc)
    out.append((name[:70], vals))
md = "--md" in sys.argv
labels = [w[0] for w in want]
if md:
    print("| kernel | " + " | ".join(labels) + " |")
    print("|---|" + "---:|" * len(labels))
    for n, v in out:
        print(f"| `{n}` | " + " | ".join(f"{x:.1f}" if x == x and abs(x) < 1e7 else "-" for x in v) + " |")
else:
    for n, v in out:
        print(f"{n:70s} " + " ".join(f"{l}={x:.1f}" for l, x in zip(labels, v)))
