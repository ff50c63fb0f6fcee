#!/usr/bin/env python
"""Turns `ncu -i X.ncu-rep --page raw --csv` into a compact per-launch table (markdown) for profiles/."""
import csv
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor(DMMA) %"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps act %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    with open(out, "w") as f:
        f.write("| # | kernel | " + " | ".join(c[1] for c in COLS) + " |\n")
        f.write("|---|---|" + "---|" * len(COLS) + "\n")
        for n, r in enumerate(rows[2:]):
            cells = []
            for m, _ in COLS:
                if m in hdr:
                    i = hdr.index(m)
                    v = r[i]
                    try:
                        v = "%.4g" % float(v)
                    except ValueError:
                        pass
                    cells.append("%s %s" % (v, units[i]) if units[i] and units[i] != "%" else v)
                else:
                    cells.append("-")
            f.write("| %d | %s | " % (n, r[ki].split("(")[0]) + " | ".join(cells) + " |\n")
    print("wrote", out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
