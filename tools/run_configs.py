"""Runs the five BASELINE.json configs on the GPU box and prints the results table of BASELINE.md §5
(markdown) plus one JSON object per config.  Usage (on a B200 box, from the repo root):

    python tools/run_configs.py [--out profiles/r02_results.md] [--skip-cfg5-full]

configs 1-2 go through the reference-facing host API (pgb_pfile_output_vcf: CPU selection + header +
GPU export to a file) on real pfile triples written to a tmp dir; their CPU column is the oracle's
export (reference-faithful I/O) with the same selections.  configs 3-5 are bench.py workloads
(device-resident + e2e through the C ABI); config 5 additionally runs all 200 000 variants through
pgb_export_gt_vcf into /dev/null (400 GB of VCF, ring-buffered) unless --skip-cfg5-full.
Nothing here reads /root/reference.
"""
from __future__ import annotations

import argparse
import ctypes
import gzip
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("oracle", "tools", os.path.join("pgen-rs_b200", "python")):
    sys.path.insert(0, os.path.join(ROOT, p))
import pgb200  # noqa: E402
import synth  # noqa: E402


def oracle():
    so = os.path.join(ROOT, "oracle", "_build", "liborc.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    lib = ctypes.CDLL(so)
    lib.orc_output_vcf.restype = ctypes.c_int
    lib.orc_output_vcf.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                   ctypes.c_char_p, ctypes.c_int]
    return lib


def sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 24), b""):
            h.update(blk)
    return h.hexdigest()


def small_config(name, prefix, sam_q, var_q, td, orc):
    out = os.path.join(td, name + ".gpu.vcf")
    pgb200.pfile_output_vcf(prefix, sam_q, var_q, out)  # warm-up (CUDA context, buffers)
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        st = pgb200.pfile_output_vcf(prefix, sam_q, var_q, out)
        dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, st)
    dt, st = best
    plan = pgb200.VcfPlan(prefix, sam_q, var_q)
    vi = plan.var_idx.astype(np.int64)
    si = plan.sam_idx.astype(np.int64)
    want = os.path.join(td, name + ".cpu.vcf")
    t0 = time.perf_counter()
    rc = orc.orc_output_vcf(prefix.encode(), vi.ctypes.data, len(vi), si.ctypes.data, len(si), want.encode(), 0)
    cpu_dt = time.perf_counter() - t0
    assert rc == 0
    exact = sha(out) == sha(want)
    g = int(st.genotypes)
    return {"config": name, "n_lines": int(st.n_lines), "kept_samples": int(st.n_kept_samples), "genotypes": g,
            "vcf_bytes": os.path.getsize(out), "gpu_wall_s": dt, "gpu_device_ms": st.device_ms,
            "gpu_export_call_ms": st.e2e_ms, "cpu_oracle_s": cpu_dt, "bit_exact": exact,
            "note": "gpu_wall_s = whole Pfile::output_vcf mirror (CPU selection + header + GPU export + file write); "
                    "cpu_oracle_s = oracle with precomputed selections (header + reference-faithful export loop)"}


def bench(workload, extra=()):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--workload", workload, "--steps", "20", "--warmup", "3", *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, check=True)
    return json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])


def cfg3_cli(td, orc):
    """configs[2] as a user runs it: `pgen-b200 filter <prefix> -o out.vcf` on a pfile triple on disk
    (lean 8-column .pvar), page cache warm, against the oracle's whole output_vcf on the same files."""
    import torch
    n, m = 2504, 1_100_000
    R = synth.record_size(n)
    prefix = os.path.join(td, "chr22")
    dev = torch.empty(m * R + 64, dtype=torch.uint8, device="cuda")
    assert pgb200.lib.pgb_dev_synth_records(dev.data_ptr(), R, 3, 0, m, n, torch.cuda.current_stream().cuda_stream) == 0
    with open(prefix + ".pgen", "wb") as f:
        f.write(synth.pgen_header(m, n))
        f.write(dev[:m * R].cpu().numpy().tobytes())
    del dev
    synth.write_pvar(prefix + ".pvar", "lean", m, 3)
    synth.write_psam(prefix + ".psam", n)
    cli = os.path.join(ROOT, "bin", "pgen-b200")
    out = os.path.join(td, "chr22.gpu.vcf")
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        subprocess.run([cli, "filter", prefix, "-o", out], check=True)
        ts.append(time.perf_counter() - t0)
    want = os.path.join(td, "chr22.cpu.vcf")
    t0 = time.perf_counter()
    rc = orc.orc_output_vcf(prefix.encode(), None, -1, None, -1, want.encode(), 0)
    cpu_dt = time.perf_counter() - t0
    assert rc == 0
    exact = sha(out) == sha(want)
    size = os.path.getsize(out)
    os.unlink(want)
    return {"config": "3 chr22 keep-all via the CLI (files in, file out)", "genotypes": n * m, "vcf_bytes": size,
            "gpu_wall_s": min(ts[1:]), "gpu_first_run_s": ts[0], "gpu_device_ms": float("nan"), "cpu_oracle_s": cpu_dt,
            "bit_exact": exact, "note": "whole process: CUDA context creation, .pvar/.psam parse, header, export, file write to "
                                        + td + " (no fsync); CPU column = oracle whole output_vcf, reference-faithful I/O"}


def cfg5_full():
    """All 200 000 variants x 500 000 samples: 25 GB page-locked .pgen image (synthesised on the device,
    copied to the host once), exported through pgb_export_gt_vcf into /dev/null."""
    import torch
    n, m = 500_000, 200_000
    R = synth.record_size(n)
    image = torch.empty(12 + m * R, dtype=torch.uint8, pin_memory=True)
    image[:12] = torch.from_numpy(np.frombuffer(synth.pgen_header(m, n), dtype=np.uint8).copy())
    blk = 8192
    dev = torch.empty(blk * R + 64, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for a in range(0, m, blk):
        k = min(blk, m - a)
        assert pgb200.lib.pgb_dev_synth_records_fast(dev.data_ptr(), R, 5, a, k, n, st) == 0
        image[12 + a * R:12 + (a + k) * R].copy_(dev[:k * R])
    torch.cuda.synchronize()
    del dev
    blob, off = synth.uniform_prefix_blob(m, 0, 40)
    fd = os.open("/dev/null", os.O_WRONLY)
    try:
        with pgb200.PgenFile(image_ptr=image.data_ptr(), image_bytes=image.numel()) as f:
            dts = []
            for _ in range(2):  # the second pass finds every buffer allocated
                t0 = time.perf_counter()
                s = f.export_gt_vcf(None, None, blob, off, fd, devices=[0])
                dts.append(time.perf_counter() - t0)
            dt = dts[1]
    finally:
        os.close(fd)
    return {"config": "5-full", "genotypes": int(s.genotypes), "vcf_bytes": int(s.bytes_out), "wall_s": dt,
            "genotypes_per_s": s.genotypes / dt, "vcf_gb_per_s": s.bytes_out / dt / 1e9, "device_ms_sum": s.device_ms,
            "device_genotypes_per_s": s.genotypes / (s.device_ms * 1e-3), "chunks": int(s.n_chunks),
            "first_pass_wall_s": dts[0], "sink": "/dev/null (ordered writes), output ring-buffered through 3 page-locked slots"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--skip-cfg5-full", action="store_true")
    ap.add_argument("--skip-cli", action="store_true")
    a = ap.parse_args()
    orc = oracle()
    res = []
    with tempfile.TemporaryDirectory(prefix="pgb_cfg_") as td:
        data = os.path.join(ROOT, "tests", "data")
        b1 = os.path.join(td, "basic1")
        open(b1 + ".pvar", "wb").write(gzip.open(os.path.join(data, "basic1.pvar.gz")).read())
        open(b1 + ".psam", "wb").write(gzip.open(os.path.join(data, "basic1.psam.txt.gz")).read())
        synth.write_pgen(b1 + ".pgen", 1, 17784, 2504)
        res.append(small_config("1 basic1 filter", b1, 'IID == "NA20900"', 'ALT == "G"', td, orc))
        r1 = os.path.join(td, "random1")
        open(r1 + ".psam", "wb").write(gzip.open(os.path.join(data, "random1.psam.txt.gz")).read())
        synth.write_pvar(r1 + ".pvar", "random1", 200000, 2)
        synth.write_pgen(r1 + ".pgen", 2, 200000, 300)
        res.append(small_config("2 random1 full", r1, None, None, td, orc))
        if not a.skip_cli:
            res.append(cfg3_cli(td, orc))
    for name, wl in (("3 chr22 keep-all", "chr22"), ("4 chr22 gather", "gather"), ("5 biobank block", "biobank-block")):
        d = bench(wl, ("--no-config5", "--no-other") if wl == "chr22" else ("--no-file", "--no-config5"))
        d["config_name"] = name
        res.append(d)
    if not a.skip_cfg5_full:
        res.append(cfg5_full())
    lines = ["| Config | genotypes | device-resident genotypes/s | K2 GB/s (frac of measured peak) | e2e genotypes/s (VCF GB/s) | CPU port genotypes/s | bit-exact |",
             "|---|---|---|---|---|---|---|"]
    for r in res:
        if "gpu_wall_s" in r:
            lines.append("| %s | %d | — (device %.3f ms) | — | %.3g (whole output_vcf incl. CPU selection: %.3f s) | %.3g (%.3f s) | %s |" % (
                r["config"], r["genotypes"], r["gpu_device_ms"], r["genotypes"] / r["gpu_wall_s"], r["gpu_wall_s"],
                r["genotypes"] / r["cpu_oracle_s"], r["cpu_oracle_s"], r["bit_exact"]))
        elif "roofline" in r:
            cpu = r.get("cpu_baseline") or {}
            lines.append("| %s | %d /step | %.4g | %.0f (%.3f) | %.4g (%.1f) | %s | parity suite |" % (
                r["config_name"], r["config"]["kept_samples"] * r["config"]["kept_variants_per_gpu"], r["value"],
                r["roofline"]["achieved"], r["roofline"]["frac"], r["e2e"]["value"], r["e2e"]["vcf_gb_per_s"],
                ("%.3g" % cpu["value"]) if cpu else "—"))
        else:
            lines.append("| 5 biobank FULL (200 000 x 500 000 -> /dev/null) | %d | %.4g (sum of device ms) | — | %.4g (%.1f), %.1f s wall | — | sampled blocks in the parity suite |" % (
                r["genotypes"], r["device_genotypes_per_s"], r["genotypes_per_s"], r["vcf_gb_per_s"], r["wall_s"]))
    text = "\n".join(lines)
    print(text)
    for r in res:
        print(json.dumps(r))
    if a.out:
        with open(a.out, "w") as f:
            f.write("# BASELINE configs on one B200 (tools/run_configs.py)\n\n" + text + "\n\n```\n" +
                    "\n".join(json.dumps(r) for r in res) + "\n```\n")


if __name__ == "__main__":
    main()
