#!/usr/bin/env python
"""soak_batch.py [TRIALS] [SEED] — randomized soak of the K2 batch path (k2_batch.cuh) on a GPU box: random sample counts,
kept-sample counts, line counts (many batches per CTA), ragged prefixes, blob and rows prefix modes, every store /
stage variant; every byte against oracle/oracle_np.py.  Test infrastructure (uses the oracle as the checker)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("oracle", "tools", os.path.join("pgen-rs_b200", "python")):
    sys.path.insert(0, os.path.join(ROOT, p))
import oracle_np as onp  # noqa: E402
import pgb200  # noqa: E402
import synth  # noqa: E402

VARIANTS = [0, 0x10020000, 0x20020000, 0x30020000, 0x40020000, 0x50020000, 0x00420000, 0x03020000, 0x10000]


def main():
    trials = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2026)
    for t in range(trials):
        n = int(rng.choice([5, 64, 100, 257, 300, 640, 1000, 2504, 4001]))
        m = int(rng.integers(1, 30000))
        mode = int(rng.integers(0, 3))
        k = None if mode == 0 else int(rng.integers(0, min(n, 1200) + 1))
        sam = None if k is None else np.sort(rng.choice(n, size=k, replace=False)).astype(np.uint32)
        recs = rng.integers(0, 256, size=(m, synth.record_size(n)), dtype=np.uint8)
        lens = rng.integers(0, int(rng.choice([1, 8, 60, 200])) + 1, size=m)
        rows = m if rng.integers(0, 2) else int(rng.integers(1, m + 1))
        var = None if rows == m else np.sort(rng.choice(m, size=rows, replace=False)).astype(np.uint32)
        vr = np.arange(m) if var is None else var
        pool = rng.integers(33, 127, size=int(lens.sum()) + 1, dtype=np.uint8)
        off_all = np.zeros(m + 1, np.uint64)
        off_all[1:] = np.cumsum(lens)
        pre = [pool[int(off_all[v]):int(off_all[v + 1])].tobytes() for v in vr]
        variant = int(rng.choice(VARIANTS))
        os.environ["PGB_K2_VARIANT"] = str(variant)
        os.environ["PGB_CHUNK_MB"] = str(int(rng.choice([1, 4, 128])))
        image = np.concatenate([np.frombuffer(synth.pgen_header(m, n), dtype=np.uint8), recs.reshape(-1)])
        with pgb200.PgenFile(image=image) as f:
            if rng.integers(0, 2):  # rows mode: prefixes = rows of a raw image + "\tGT" appended on the device
                got = pgb200.export_rows_to_bytes(f, var, sam, pool, off_all[vr], lens[vr].astype(np.uint32))
                want = onp.format_body(recs, vr, np.arange(n) if sam is None else sam, [p + b"\tGT" for p in pre])
            else:
                blob = np.frombuffer(b"".join(pre) + b"\0", dtype=np.uint8).copy()
                off = np.zeros(len(vr) + 1, np.uint64)
                off[1:] = np.cumsum([len(p) for p in pre])
                got = pgb200.export_to_bytes(f, var, sam, blob, off)
                want = onp.format_body(recs, vr, np.arange(n) if sam is None else sam, pre)
        if got != want:
            print("MISMATCH trial %d: n=%d m=%d rows=%d k=%s variant=%#x" % (t, n, m, len(vr), k, variant))
            return 1
    print("soak ok: %d trials" % trials)
    return 0


if __name__ == "__main__":
    sys.exit(main())
