"""Writes a chr22-shaped pfile triple to a directory and runs bin/pgen-b200 filter on it with PGB_TRACE=1."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tools", os.path.join("pgen-rs_b200", "python")):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
import pgb200, synth

td = sys.argv[1] if len(sys.argv) > 1 else "/tmp/pgb_cli"
os.makedirs(td, exist_ok=True)
n, m = 2504, int(os.environ.get("M", "1100000"))
R = synth.record_size(n)
prefix = os.path.join(td, "chr22")
dev = torch.empty(m * R + 64, dtype=torch.uint8, device="cuda")
assert pgb200.lib.pgb_dev_synth_records(dev.data_ptr(), R, 3, 0, m, n, torch.cuda.current_stream().cuda_stream) == 0
with open(prefix + ".pgen", "wb") as f:
    f.write(synth.pgen_header(m, n))
    f.write(dev[:m * R].cpu().numpy().tobytes())
del dev
torch.cuda.empty_cache()
synth.write_pvar(prefix + ".pvar", "lean", m, 3)
synth.write_psam(prefix + ".psam", n)
cli = os.path.join(ROOT, "bin", "pgen-b200")
for out in (os.path.join(td, "out.vcf"), "/dev/shm/pgb_cli_out.vcf", "/dev/null"):
    env = dict(os.environ, PGB_TRACE="1")
    t0 = time.perf_counter()
    r = subprocess.run([cli, "filter", prefix, "-o", out], env=env, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    lines = r.stderr.splitlines()
    keep = [l for l in lines if "chunk" not in l] + [l for l in lines if "chunk 0 " in l or "chunk 41 " in l]
    print(f"== -o {out}: {dt:.2f} s, rc={r.returncode}")
    print("\n".join(keep[-24:]))
    if out not in ("/dev/null",) and os.path.exists(out):
        os.unlink(out)
