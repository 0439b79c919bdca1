"""Store-only calibration sweep: plain grid-stride fill vs K2-like write patterns
(pgb_dev_fill_pattern).  Run on the GPU box: python tools/fill_sweep.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pgen-rs_b200", "python"))
import torch
import pgb200

lib = pgb200.lib
n = 11_062_700_000 // 16 * 16
buf = torch.empty(n + 1024, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream


def run(fn, reps=5):
    for _ in range(2):
        assert fn() == 0
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return reps * n / (a.elapsed_time(b) * 1e-3) / 1e9


print(f"pgb_dev_fill: {run(lambda: lib.pgb_dev_fill(buf.data_ptr(), n, 0, st)):8.1f} GB/s")
for ctas in (148 * 4, 148 * 6, 148 * 8, 148 * 16, 148 * 32):
    for burst in (1, 2, 4):
        r = run(lambda: lib.pgb_dev_fill_pattern(buf.data_ptr(), n, 0, 1, burst, ctas, 0, st))
        print(f"strided ctas={ctas:5d} burst={burst}: {r:8.1f} GB/s")
for rows in (1, 2, 4, 8, 20):
    for coop in (1, 8):
        for burst in (1, 2, 4):
            if burst > rows:
                continue
            for ctas in (148 * 2, 148 * 6, 148 * 8, 148 * 32):
                r = run(lambda: lib.pgb_dev_fill_pattern(buf.data_ptr(), n, rows, coop, burst, ctas, 0, st))
                print(f"seg rows={rows:3d} coop={coop} burst={burst} ctas={ctas:5d}: {r:8.1f} GB/s")
