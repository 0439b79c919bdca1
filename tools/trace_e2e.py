"""PGB_TRACE timeline of one warm export of a bench workload through the C ABI (page-locked in/out)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tools", os.path.join("pgen-rs_b200", "python"), ROOT):
    sys.path.insert(0, os.path.join(ROOT, p) if not os.path.isabs(p) else p)
import numpy as np
import torch
import pgb200, synth
from bench import WORKLOADS

wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "gather"]
n, m, width = wl["n"], wl["m"], wl["width"]
R = synth.record_size(n)
dev = torch.empty(m * R + 64, dtype=torch.uint8, device="cuda")
assert pgb200.lib.pgb_dev_synth_records(dev.data_ptr(), R, wl["seed"], 0, m, n, torch.cuda.current_stream().cuda_stream) == 0
image = torch.empty(12 + m * R, dtype=torch.uint8, pin_memory=True)
image[:12] = torch.from_numpy(np.frombuffer(synth.pgen_header(m, n), dtype=np.uint8).copy())
image[12:].copy_(dev[:m * R])
torch.cuda.synchronize()
var = None if wl["mk"] is None else synth.subset_indices(42, m, wl["mk"])
sam = None if wl["k"] is None else synth.subset_indices(41, n, wl["k"])
b, _ = synth.uniform_prefix_blob(m, 0, width)
if var is not None:
    b = np.ascontiguousarray(b.reshape(m, width)[var]).reshape(-1)
nl = m if var is None else len(var)
off = np.arange(nl + 1, dtype=np.uint64) * np.uint64(width)
K = n if sam is None else len(sam)
total = int(off[-1]) + nl * (4 * K + 1)
out = torch.empty(total + 64, dtype=torch.uint8, pin_memory=True)
with pgb200.PgenFile(image_ptr=image.data_ptr(), image_bytes=image.numel()) as f:
    for i in range(3):
        if i == 2:
            os.environ["PGB_TRACE"] = "1"
        t0 = time.perf_counter()
        _, st = f.export_gt_vcf_mem(var, sam, b, off, out.data_ptr(), total, devices=[0])
        print("call %d: %.2f ms, chunks %d" % (i, 1e3 * (time.perf_counter() - t0), st.n_chunks), flush=True)
