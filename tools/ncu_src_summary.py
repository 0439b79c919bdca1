#!/usr/bin/env python
"""ncu_src_summary.py REPORT.csv — summarise `ncu --page source --csv` (SASS view): instructions
executed and stall samples per source line / per SASS instruction, top N."""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = list(csv.reader(open(path)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {n: i for i, n in enumerate(hdr)}
    tot_inst = tot_samp = 0
    items = []
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr):
            continue
        try:
            inst = int(r[col["Instructions Executed"]] or 0)
            samp = int(r[col["# Samples"]] or 0)
        except ValueError:
            continue
        tot_inst += inst
        tot_samp += samp
        items.append((r[col["Address"]], r[col["Source"]], inst, samp, r))
    print(f"total warp instructions {tot_inst}, samples {tot_samp}")
    stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    agg = defaultdict(int)
    for _, _, _, _, r in items:
        for n in stall_cols:
            try:
                agg[n] += int(r[col[n]] or 0)
            except ValueError:
                pass
    print("stall samples:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
    print("\n-- all instructions in address order (inst, samples) --")
    for a, s, inst, samp, r in items:
        if inst * 400 >= tot_inst or samp * 200 >= tot_samp:
            extra = ""
            if "L1 Wavefronts Shared" in col and r[col["L1 Wavefronts Shared"]] not in ("", "0"):
                extra = f" smem_wf={r[col['L1 Wavefronts Shared']]} ideal={r[col['L1 Wavefronts Shared Ideal']]}"
            print(f"{a[-5:]} {inst:>10} {samp:>6}  {s[:90]}{extra}")


if __name__ == "__main__":
    main()
