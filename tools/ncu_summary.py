"""One markdown row per ncu report: duration, DRAM traffic, pipe utilisation, occupancy (reads `--page raw --csv`).

    python tools/ncu_summary.py gpurun_out/prof_*.ncu-rep
"""
import csv
import json
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__inst_executed_pipe_tensor_op_gmma.avg.pct_of_peak_sustained_active", "tcgen05 %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v) * mult.get(unit, 1)


def main():
    out = {}
    print("| kernel | " + " | ".join(n for _, n in WANT) + " |")
    print("|---|" + "---|" * len(WANT))
    for rep in sys.argv[1:]:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units, vals = rows[0], rows[1], rows[-1]
        idx = {k: i for i, k in enumerate(hdr)}
        name = vals[idx["Kernel Name"]].split("(")[0]
        cells = []
        rec = {}
        for key, label in WANT:
            if key not in idx:
                cells.append("-")
                continue
            v, u = vals[idx[key]], units[idx[key]]
            rec[key] = (v, u)
            try:
                f = float(v)
                cells.append(f"{f:.4g} {u}".strip())
            except ValueError:
                cells.append(v)
        print(f"| `{name}` | " + " | ".join(cells) + " |")
        rd = rec.get("dram__bytes_read.sum")
        wr = rec.get("dram__bytes_write.sum")
        if rd and wr:
            out[name] = int(to_bytes(*rd) + to_bytes(*wr))
    print()
    print("dram bytes per launch:", json.dumps(out))


if __name__ == "__main__":
    main()
