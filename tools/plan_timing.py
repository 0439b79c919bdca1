#!/usr/bin/env python
"""plan_timing.py [ROWS] [KIND] — CPU planning stage of output_vcf on a synthetic .pvar (default 1 100 000 rows of
the '1000g' kind, ~170-byte rows): time of pgb_plan_vcf (selection + header + row positions; what output_vcf now
does before the export) and of building the prefix blob on top (what it did until round 2; now only on demand).
No GPU needed."""
import ctypes as C
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tools"), os.path.join(ROOT, "pgen-rs_b200", "python")]
import pgb200  # noqa: E402
import synth  # noqa: E402


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_100_000
    kind = sys.argv[2] if len(sys.argv) > 2 else "1000g"
    with tempfile.TemporaryDirectory(prefix="pgb_plan_") as td:
        prefix = os.path.join(td, "t")
        synth.write_pvar(prefix + ".pvar", kind, rows, 3)
        synth.write_psam(prefix + ".psam", 2504)
        with open(prefix + ".pgen", "wb") as f:
            f.write(synth.pgen_header(rows, 2504))
        size = os.path.getsize(prefix + ".pvar")
        lib = pgb200.lib
        for query in (None, 'FILTER == "PASS"'):
            best_plan = best_blob = 1e9
            for _ in range(3):
                h = C.c_void_p()
                t0 = time.perf_counter()
                rc = lib.pgb_plan_vcf(prefix.encode(), None, None if query is None else query.encode(), C.byref(h))
                t1 = time.perf_counter()
                assert rc == 0
                ln = C.c_uint64()
                lib.pgb_plan_prefix_blob(h, C.byref(ln))  # materialises the blob (the former per-row host pass)
                t2 = time.perf_counter()
                nv = lib.pgb_plan_n_var(h)
                lib.pgb_plan_free(h)
                best_plan, best_blob = min(best_plan, t1 - t0), min(best_blob, t2 - t1)
            print(f"{rows} rows ({size / 1e6:.0f} MB .pvar, kind {kind}), var query {query!r}: kept {nv}; "
                  f"plan (rows handed to the device) {best_plan * 1e3:.0f} ms; + host prefix blob ({ln.value / 1e6:.0f} MB) "
                  f"{best_blob * 1e3:.0f} ms -> before {1e3 * (best_plan + best_blob):.0f} ms, after {best_plan * 1e3:.0f} ms "
                  f"on {os.cpu_count()} cores")


if __name__ == "__main__":
    main()
