"""Small-shape driver for compute-sanitizer (memcheck / racecheck): a few exports through the C
ABI covering keep-all, gather, empty selections, multi-tile lines and every K2 tuning variant,
each checked against the numpy oracle.  Usage on the GPU box:
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("oracle", "tools", os.path.join("pgen-rs_b200", "python")):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np

import oracle_np as onp
import pgb200
import synth


def image_of(recs, n):
    return np.concatenate([np.frombuffer(synth.pgen_header(recs.shape[0], n), dtype=np.uint8), recs.reshape(-1)])


def main():
    rng = np.random.default_rng(123)
    cases = 0
    for variant in (0x000, 0x010, 0x102, 0x812, 0x1200):
        os.environ["PGB_K2_VARIANT"] = str(variant)
        for n, m in ((1, 3), (5, 7), (301, 40), (2504, 33), (20000, 5)):
            recs = rng.integers(0, 256, size=(m, synth.record_size(n)), dtype=np.uint8)
            for sam in (None, np.zeros(0, np.uint32), np.sort(rng.choice(n, size=max(1, n // 3), replace=False)).astype(np.uint32)):
                var = np.sort(rng.choice(m, size=max(1, m // 2), replace=False)).astype(np.uint32)
                pre = [bytes(rng.integers(33, 127, size=rng.integers(0, 90), dtype=np.uint8)) for _ in var]
                blob = np.frombuffer(b"".join(pre) + b"\0", dtype=np.uint8).copy()
                off = np.zeros(len(var) + 1, np.uint64)
                off[1:] = np.cumsum([len(x) for x in pre])
                with pgb200.PgenFile(image=image_of(recs, n)) as f:
                    got = pgb200.export_to_bytes(f, var, sam, blob, off)
                want = onp.format_body(recs, var, np.arange(n) if sam is None else sam, pre)
                assert got == want, (variant, n, m)
                cases += 1
    print("sanitize_small: %d cases bit-exact" % cases)


if __name__ == "__main__":
    main()
