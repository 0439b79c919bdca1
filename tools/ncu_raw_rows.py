#!/usr/bin/env python
"""ncu_raw_rows.py REPORT.ncu-rep [> profiles/NAME_raw.txt] — the rows of `ncu --page raw` that the rooflines in
DESIGN.md / bench.py are computed from (per launch in the report): duration, DRAM bytes read/written, DRAM / L1 /
L2 / issue utilisation, executed instructions, registers, grid.  Runs here (no GPU needed)."""
import csv
import subprocess
import sys

KEEP = (
    "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
)


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        for h, u, v in zip(hdr, units, vals):
            if h in KEEP:
                print(f"{h:66s} {u:16s} {v}")
        print()


if __name__ == "__main__":
    main()
