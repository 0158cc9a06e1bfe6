"""Per-source-line summary of an ncu report's source page (needs -lineinfo and --import-source on).

    python tools/ncu_lines.py <report.ncu-rep> [top_n] [file_filter]
Prints, per CUDA source line: stall samples, instructions executed, shared-memory wavefronts (actual / ideal).
"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    filt = sys.argv[3] if len(sys.argv) > 3 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, hdr, lines = None, None, []
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1]
            continue
        if len(r) >= 2 and r[0] == "Function Name":
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or not r:
            continue
        if r[0] != "":  # a CUDA source line row (aggregated over its SASS)
            d = dict(zip(hdr, r))
            def num(k):
                try:
                    return float(d.get(k, "0") or 0)
                except ValueError:
                    return 0.0
            lines.append((cur_file, int(r[0]), r[1].strip(), num("# Samples"), num("Instructions Executed"),
                          num("L1 Wavefronts Shared"), num("L1 Wavefronts Shared Ideal"), num("stall_barrier"),
                          num("stall_long_sb"), num("stall_short_sb"), num("stall_mio"), num("stall_wait"), num("stall_math")))
    tot_s = sum(l[3] for l in lines) or 1
    tot_i = sum(l[4] for l in lines) or 1
    print(f"total samples {tot_s:.0f}, instructions {tot_i:.0f}")
    lines = [l for l in lines if filt in l[0]]
    lines.sort(key=lambda l: -l[3])
    print("samples%  inst%   shWave/ideal    bar  lsb  ssb  mio wait math  file:line  source")
    for f, ln, src, smp, ins, w, wi, sb, slb, ssb, smio, sw, sm in lines[:top]:
        print(f"{100 * smp / tot_s:6.2f} {100 * ins / tot_i:6.2f}  {w / 1e6:7.2f}/{wi / 1e6:<7.2f} {sb:5.0f}{slb:5.0f}{ssb:5.0f}{smio:5.0f}{sw:5.0f}{sm:5.0f}"
              f"  {f.split('/')[-1]}:{ln}  {src[:90]}")


if __name__ == "__main__":
    main()
