"""Times one export of a biobank-shaped block through the C ABI into different sinks
(page-locked memory, pageable memory, /dev/null, tmpfs file) to separate PCIe from sink cost."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tools", os.path.join("pgen-rs_b200", "python")):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
import torch
import pgb200, synth

n, m = 500_000, int(os.environ.get("M", "8192"))
R = synth.record_size(n)
dev = torch.empty(m * R + 64, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
assert pgb200.lib.pgb_dev_synth_records(dev.data_ptr(), R, 5, 0, m, n, st) == 0
image = torch.empty(12 + m * R, dtype=torch.uint8, pin_memory=True)
image[:12] = torch.from_numpy(np.frombuffer(synth.pgen_header(m, n), dtype=np.uint8).copy())
image[12:].copy_(dev[:m * R])
torch.cuda.synchronize()
blob, off = synth.uniform_prefix_blob(m, 0, 40)
total = int(off[-1]) + m * (4 * n + 1)
pinned = torch.empty(total + 64, dtype=torch.uint8, pin_memory=True)
pageable = np.empty(total + 64, dtype=np.uint8)
with pgb200.PgenFile(image_ptr=image.data_ptr(), image_bytes=image.numel()) as f:
    def run(name, fn, reps=3):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            s = fn()
        dt = (time.perf_counter() - t0) / reps
        print(f"{name:28s} {total / dt / 1e9:7.2f} GB/s  {dt * 1e3:8.1f} ms  chunks={s.n_chunks} device_ms={s.device_ms:.1f}", flush=True)
    run("pinned memory", lambda: f.export_gt_vcf_mem(None, None, blob, off, pinned.data_ptr(), total, devices=[0])[1])
    run("pageable memory", lambda: f.export_gt_vcf_mem(None, None, blob, off, pageable.ctypes.data, total, devices=[0])[1])
    for w in ("1", "8"):
        os.environ["PGB_WRITERS"] = w
        fd = os.open("/dev/null", os.O_WRONLY)
        run(f"/dev/null writers={w}", lambda: f.export_gt_vcf(None, None, blob, off, fd, devices=[0]))
        os.close(fd)
    for mb in ("64", "1024"):
        os.environ["PGB_CHUNK_MB"] = mb
        fd = os.open("/dev/null", os.O_WRONLY)
        run(f"/dev/null chunk={mb}MB", lambda: f.export_gt_vcf(None, None, blob, off, fd, devices=[0]))
        os.close(fd)
    os.environ["PGB_CHUNK_MB"] = "256"
    def tofile():
        fd = os.open("/dev/shm/pgb_sink.vcf", os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
        try:
            return f.export_gt_vcf(None, None, blob, off, fd, devices=[0])
        finally:
            os.close(fd)
    run("tmpfs file writers=8", tofile, reps=2)
    os.unlink("/dev/shm/pgb_sink.vcf")
