#!/usr/bin/env python
"""bench.py — throughput of the genotype export hot path (decode 2-bit hardcalls -> gather
kept samples -> VCF GT text) on B200, with the CPU restatement of pgen-rs timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

One "step" = one pass of the hot path (K1 record index + line-length prefix sum, K2 decode +
gather + format) over one batch of synthetic input:

  workload chr22  (default; BASELINE.json configs[2], the largest configuration whose whole
                  output is HBM-resident on one GPU): 2 504 samples x 1 100 000 variants,
                  keep all, 40-byte `.pvar` prefixes -> 11.06 GB of VCF body per step.
  workload gather (configs[3]): same matrix, 250 of 2 504 samples, 550 000 of 1 100 000 variants.
  workload biobank-block (configs[4] shape): 500 000 samples x 8 192-variant block.

`value`  = genotypes decoded+emitted per second, inputs and output resident in HBM, CUDA
           events on the launching stream, max over ranks.
`e2e`    = the same metric through the reference-facing C ABI call (pgb_export_gt_vcf_mem)
           with HOST buffers: page-locked .pgen image in, page-locked VCF body out, H2D and
           D2H inside the timed region.
Multi-GPU (torchrun): rank r owns the contiguous variant range [r*M, (r+1)*M) of an
(N_gpus*M)-variant matrix — weak scaling, no data-path collective (NCCL only for the timing
barrier and the max-over-ranks reduction).

The default line also carries `other_workloads` (device-resident K1 + K2 of the gather, random1 and biobank-block
shapes: the rooflines of the kernels chr22 does not exercise) and `config5`: BASELINE.json configs[4] (500 000 samples x 200 000 variants, 400 GB of
VCF) in STRONG scaling — rank r formats the contiguous variant range [r*200000/N, (r+1)*200000/N), records resident
in HBM, output ring-buffered — plus one call of the product's own multi-device sharding
(pgb_export_gt_vcf_mem(devices=[0..N-1]) from rank 0) on a slice of that matrix that crosses the 4 GiB record
offset, checked bit for bit against the device-resident path.

`--impl reference` times the CPU restatement of Pfile::output_vcf (oracle/pgen_oracle.c in its
reference-faithful I/O mode: lseek+read per variant, 8 KiB buffered writer, two appends per
genotype; the Rust binary itself cannot be built in this image) on a bounded sample of the
same workload, single-threaded like pgen-rs.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "tools"), os.path.join(ROOT, "pgen-rs_b200", "python")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import synth  # noqa: E402

METRIC = "genotypes decoded+emitted/sec (var x kept-sample)"
UNIT = "genotypes/s"

WORKLOADS = {
    # name: (n_samples, n_variants per GPU, kept samples or None, kept variants or None, prefix width, seed)
    "chr22": dict(n=2504, m=1_100_000, k=None, mk=None, width=40, seed=3,
                  desc="configs[2]: synthetic 1000G chr22 shape, 2504 samples x 1.1M variants, keep all"),
    "chr22-wide-prefix": dict(n=2504, m=1_100_000, k=None, mk=None, width=163, seed=3,
                              desc="configs[2] shape with 163-byte line prefixes (the mean of the real 1000G basic1.pvar)"),
    "random1": dict(n=300, m=200_000, k=None, mk=None, width=40, seed=2,
                    desc="configs[1] shape: 300 samples x 200000 variants, keep all (short 1.2 KB lines)"),
    "gather": dict(n=2504, m=1_100_000, k=250, mk=550_000, width=40, seed=3,
                   desc="configs[3]: chr22 shape, 10% samples (250) x 50% variants (550000)"),
    "biobank-block": dict(n=500_000, m=8192, k=None, mk=None, width=40, seed=5,
                          desc="configs[4] shape: 500000 samples x 8192-variant block of the 200000"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------- clocks ---
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU during the timed regions (NVML)."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._th = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                try:
                    idx = int(vis.split(",")[self.index])
                except ValueError:
                    pass
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        except Exception:
            return self
        names = {
            "hw_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(pynvml, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(pynvml, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }

        def loop():
            while not self._stop.is_set():
                try:
                    mhz = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                    util = pynvml.nvmlDeviceGetUtilizationRates(h).gpu
                    try:
                        r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                    except Exception:
                        r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    self.samples.append((mhz, util))
                    for k, bit in names.items():
                        if r & bit:
                            self.reasons.add(k)
                except Exception:
                    pass
                self._stop.wait(0.02)

        self._th = threading.Thread(target=loop, daemon=True)
        self._th.start()
        return self

    def stop(self):
        self._stop.set()
        if self._th:
            self._th.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(s[0] for s in self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------- CPU reference ---
def _load_oracle():
    so = os.path.join(ROOT, "oracle", "_build", "liborc.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    lib = ctypes.CDLL(so)
    lib.orc_export_body.restype = ctypes.c_int
    lib.orc_export_body.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    return lib


class CpuReference:
    """The oracle in reference-faithful I/O mode on the first `rows` variants of the workload
    (same seed, same prefixes, same sample selection), writing a real file like pfile.rs:136."""

    def __init__(self, wl, tmpdir):
        self.wl = wl
        self.lib = _load_oracle()
        self.tmpdir = tmpdir
        self.sam = None if wl["k"] is None else synth.subset_indices(41, wl["n"], wl["k"])
        self.rows_on_disk = 0
        self.pgen = os.path.join(tmpdir, "sample.pgen")
        self.out = os.path.join(tmpdir, "sample.vcf")

    def prepare(self, rows):
        wl = self.wl
        if rows <= self.rows_on_disk:
            return
        # the header claims the full workload's M; only the first `rows` records are present
        with open(self.pgen, "wb") as f:
            f.write(synth.pgen_header(wl["m"], wl["n"]))
            step = max(1, (64 << 20) // synth.record_size(wl["n"]))
            for a in range(0, rows, step):
                f.write(synth.synth_records(wl["seed"], a, min(step, rows - a), wl["n"]).tobytes())
        self.rows_on_disk = rows
        if wl["mk"] is None:
            self.var_all = np.arange(rows, dtype=np.uint32)
        else:
            full = synth.subset_indices(42, wl["m"], wl["mk"])
            self.var_all = full[full < rows]
        blob, _ = synth.uniform_prefix_blob(rows, 0, wl["width"])
        self.blob_rows = blob.reshape(rows, wl["width"])

    def run(self, rows):
        """Returns (seconds, genotypes) for the kept variants among file rows [0, rows)."""
        wl = self.wl
        self.prepare(rows)
        var = self.var_all[self.var_all < rows]
        blob = np.ascontiguousarray(self.blob_rows[var]).reshape(-1)
        off = np.arange(len(var) + 1, dtype=np.uint64) * np.uint64(wl["width"])
        k = wl["n"] if self.sam is None else len(self.sam)
        fd = os.open(self.out, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)  # File::create, pfile.rs:136
        try:
            t0 = time.perf_counter()
            rc = self.lib.orc_export_body(self.pgen.encode(), var.ctypes.data, len(var),
                                          None if self.sam is None else self.sam.ctypes.data, k,
                                          blob.ctypes.data, off.ctypes.data, fd, 0)
            dt = time.perf_counter() - t0
        finally:
            os.close(fd)
        if rc != 0:
            raise RuntimeError(f"oracle failed: {rc}")
        return dt, len(var) * k

    def rows_for_seconds(self, seconds):
        """Calibrate on a block of >= 60 M genotypes, then size a sample worth ~`seconds` of CPU
        work, capped so that one step writes at most ~3 GB of VCF to the tmp file."""
        wl = self.wl
        k = wl["n"] if self.sam is None else len(self.sam)
        frac = 1.0 if wl["mk"] is None else wl["mk"] / wl["m"]
        probe = int(max(64, min(wl["m"], 60_000_000 / max(1.0, k * frac))))
        self.run(probe)  # page in
        dt, _ = self.run(probe)
        rate_rows = probe / max(dt, 1e-6)
        cap = int(3e9 / ((4 * k + wl["width"] + 1) * frac))
        return int(max(min(probe, cap), min(wl["m"], rate_rows * seconds, cap)))


def run_reference(args, wl, rank, world):
    if rank != 0:
        return
    # seconds of CPU work for the whole --steps/--warmup run
    budget = float(os.environ.get("PGB_BENCH_REF_BUDGET_S", "150"))
    with tempfile.TemporaryDirectory(prefix="pgb_ref_") as td:
        ref = CpuReference(wl, td)
        rows = ref.rows_for_seconds(budget / (args.steps + args.warmup))
        for _ in range(args.warmup):
            ref.run(rows)
        t = 0.0
        g = 0
        for _ in range(args.steps):
            dt, gg = ref.run(rows)
            t += dt
            g += gg
    value = g / t
    sample = f"first {rows} of {wl['m']} variants per step ({g // args.steps} genotypes/step), reference-faithful I/O, output to a tmp file"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": args.workload, "description": wl["desc"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                         "note": "C restatement of pgen-rs Pfile::output_vcf (single-threaded like the Rust original, which cannot be built here)",
                         "host_cores": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)



# ------------------------------------------------- host-side ceilings (measured, per run) ---
def host_write_probe(h_out_np, n_threads):
    """CPU threads filling the page-locked output buffer (numpy copies release the GIL): what the host
    memory system takes when the writers are cores instead of the GPUs' DMA engines."""
    nbytes = min(h_out_np.nbytes, 4 << 30)
    per = nbytes // n_threads // 4096 * 4096
    src = np.full(per, 48, np.uint8)

    def work(i):
        np.copyto(h_out_np[i * per:(i + 1) * per], src)

    best = 0.0
    for _ in range(2):
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(i,)) for i in range(n_threads)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        best = max(best, per * n_threads / (time.perf_counter() - t0) / 1e9)
    return best


def d2h_calibration(torch, d_out, h_out, total, barrier):
    """Device->host ceilings of this rank, all ranks copying at the same time (the host side is shared):
    (a) ONE copy of the whole body, (b) the export's own pattern without kernels — 128 MiB chunks cycling
    over three streams.  The e2e roofline peak is the larger of the two."""
    ca, cb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h_out[:1 << 20].copy_(d_out[:1 << 20], non_blocking=True)
    barrier()
    ca.record()
    h_out[:total].copy_(d_out[:total], non_blocking=True)
    cb.record()
    torch.cuda.synchronize()
    whole_s = ca.elapsed_time(cb) * 1e-3
    whole = total / whole_s / 1e9
    streams = [torch.cuda.Stream() for _ in range(3)]
    chunk = 128 << 20
    barrier()
    t0 = time.perf_counter()
    for i, a in enumerate(range(0, total, chunk)):
        b = min(total, a + chunk)
        with torch.cuda.stream(streams[i % 3]):
            h_out[a:b].copy_(d_out[a:b], non_blocking=True)
    torch.cuda.synchronize()
    chunked_s = time.perf_counter() - t0
    chunked = total / chunked_s / 1e9
    return {"whole_body_gbs": whole, "chunked_3_streams_128MiB_gbs": chunked, "peak": max(whole, chunked),
            "whole_body_s": whole_s, "chunked_s": chunked_s}


def file_sink_ceilings(path, mv, n_threads):
    """Raw rates of the candidate sink strategies on the file system of `path`, for the same bytes the export
    writes (no GPU involved): n_threads buffered pwrite()s, n_threads O_DIRECT pwrite()s, n_threads copies
    through a MAP_SHARED mapping, and one buffered pwrite() thread.  No fsync, like the export."""
    import mmap
    total = len(mv) // 4096 * 4096
    per = total // n_threads // 4096 * 4096
    out = {}

    def run(name, flags, nt, use_map=False):
        try:
            fd = os.open(path, os.O_RDWR | os.O_CREAT | os.O_TRUNC | flags, 0o644)
        except OSError as e:
            out[name] = {"error": str(e)}
            return
        try:
            os.ftruncate(fd, total)
            m = mmap.mmap(fd, total) if use_map else None
            mnp = np.frombuffer(m, dtype=np.uint8) if use_map else None
            srcnp = np.frombuffer(mv, dtype=np.uint8)
            span = total // nt // 4096 * 4096
            err = []

            def work(t):
                try:
                    a, e = t * span, (t + 1) * span
                    if use_map:
                        np.copyto(mnp[a:e], srcnp[a:e])
                        return
                    o = a
                    while o < e:
                        o += os.pwrite(fd, mv[o:min(e, o + (64 << 20))], o)
                except OSError as ex:
                    err.append(str(ex))

            t0 = time.perf_counter()
            th = [threading.Thread(target=work, args=(t,)) for t in range(nt)]
            for x in th:
                x.start()
            for x in th:
                x.join()
            dt = time.perf_counter() - t0
            out[name] = {"error": err[0]} if err else span * nt / dt / 1e9
            if use_map:
                del mnp
                m.close()
        finally:
            os.close(fd)
            try:
                os.unlink(path)
            except OSError:
                pass

    run("buffered_pwrite_1", 0, 1)
    run("buffered_pwrite_%d" % n_threads, 0, n_threads)
    run("o_direct_pwrite_%d" % n_threads, os.O_DIRECT, n_threads)
    run("mmap_copy_%d" % n_threads, 0, n_threads, use_map=True)
    del per
    return out


# --------------------------------------------------------- config 5, strong scaling ---
C5 = dict(n=500_000, m=200_000, width=40, seed=5, block=2048)
C5_N1_FILE = os.path.join(tempfile.gettempdir(), "pgb_config5_n1.json")


def run_config5(args, torch, pgb200, lib, dev, rank, world, barrier, cpu_barrier, max_over_ranks):
    """BASELINE.json configs[4]: 500 000 samples x 200 000 variants (25 GB of records, 400 GB of VCF).
    Strong scaling: rank r owns variants [r*M/N, (r+1)*M/N); its records are generated on the device and stay
    resident; K1 + K2 run block by block (2 048 variants = 4.1 GB of text) into a two-buffer output ring.
    Then ONE call of the product's multi-device path from rank 0 on a slice crossing the 4 GiB record offset."""
    n, m_all, width, seed, blk = C5["n"], args.config5_variants or C5["m"], C5["width"], C5["seed"], C5["block"]
    R = synth.record_size(n)
    L = width + 4 * n + 1
    lo, hi = rank * m_all // world, (rank + 1) * m_all // world
    rows = hi - lo
    stream = torch.cuda.current_stream().cuda_stream
    recs = torch.empty(rows * R + 64, dtype=torch.uint8, device=dev)
    recs[rows * R:].zero_()
    step = 16384
    for a in range(0, rows, step):
        pgb200._check(lib.pgb_dev_synth_records_fast(recs.data_ptr() + a * R, R, seed, lo + a, min(step, rows - a), n, stream), "synth")
    blob_np, off_np = synth.uniform_prefix_blob(rows, lo, width)
    d_blob = torch.from_numpy(blob_np).to(dev)
    d_off = torch.from_numpy(off_np.view(np.int64)).to(dev)
    ring = [torch.empty(blk * L + 1024, dtype=torch.uint8, device=dev) for _ in range(2)]
    d_meta = torch.zeros((blk + 1) * 4, dtype=torch.int64, device=dev)
    d_scr = torch.zeros(lib.pgb_dev_index_scratch_bytes(blk) // 8 + 1, dtype=torch.int64, device=dev)
    variant = int(os.environ.get("PGB_K2_VARIANT", "0"), 0)

    def block(a, out):
        nb = min(blk, rows - a)
        pgb200._check(lib.pgb_dev_index_lines(None, d_off.data_ptr() + 8 * a, 0, nb, n, R, d_meta.data_ptr(),
                                              d_scr.data_ptr(), stream), "K1")
        pgb200._check(lib.pgb_dev_format_lines_ex(recs.data_ptr() + a * R, R, d_meta.data_ptr(), nb, d_blob.data_ptr(), 0, 0,
                                                  None, n, width, out.data_ptr(), variant, stream), "K2")
        return nb

    for i in range(2):  # warm-up
        if rows:
            block(0, ring[i])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    launches = 0
    for i, a in enumerate(range(0, rows, blk)):
        block(a, ring[i & 1])
        launches += 4
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    # round trip of the last block on the device: text -> codes -> packed bytes == the records
    a = (rows - 1) // blk * blk
    nb = block(a, ring[0])
    torch.cuda.synchronize()
    lines = ring[0][:nb * L].view(nb, L)[:min(nb, 64)]
    gt = lines[:, width:width + 4 * n].reshape(-1, n, 4)
    lut = torch.full((256, 256), 255, dtype=torch.uint8, device=dev)
    for code, txt in enumerate([b"00", b"01", b"11", b".."]):
        lut[txt[0], txt[1]] = code
    codes = lut[gt[:, :, 1].long(), gt[:, :, 3].long()]
    shifts = torch.tensor([0, 2, 4, 6], dtype=torch.uint8, device=dev)
    packed = (codes.view(-1, n // 4, 4) << shifts).sum(dim=2, dtype=torch.uint8)
    rt_ok = bool(torch.equal(packed, recs[a * R:(a + lines.shape[0]) * R].view(-1, R))) and \
        bool((gt[:, :, 0] == 9).all()) and bool((lines[:, L - 1] == 10).all()) and \
        bool(torch.equal(lines[:, :width], d_blob[a * width:(a + lines.shape[0]) * width].view(-1, width)))
    del lines, gt, codes, packed, lut
    genotypes = m_all * n
    value = genotypes / (ms * 1e-3)
    alg = m_all * (R + width + L)
    peak, _ = peaks()
    rec = {"workload": "configs[4]: 500000 samples x %d variants, keep all" % m_all, "scaling": "strong",
           "variants_per_gpu": rows, "value": value, "unit": UNIT, "ms": ms, "vcf_gb_per_s": m_all * L / (ms * 1e-3) / 1e9,
           "hbm_gbs_per_gpu": alg / world / (ms * 1e-3) / 1e9, "frac_of_measured_hbm_peak": alg / world / (ms * 1e-3) / 1e9 / peak,
           "gpu_launches": launches, "block_variants": blk, "output": "two-buffer ring of %.1f GB blocks (400 GB never resident)" % (blk * L / 1e9),
           "records": "generated on the device (pgb_dev_synth_records_fast), resident in HBM: %.1f GB per GPU" % (rows * R / 1e9),
           "round_trip_ok": rt_ok, "efficiency_vs_n1": None}
    if rank == 0:
        try:
            if world == 1 and not args.config5_variants:
                with open(C5_N1_FILE, "w") as f:
                    json.dump({"value": value, "when": time.time()}, f)
            elif os.path.exists(C5_N1_FILE) and not args.config5_variants:
                with open(C5_N1_FILE) as f:
                    n1 = json.load(f)
                if time.time() - n1["when"] < 6 * 3600:
                    rec["efficiency_vs_n1"] = value / n1["value"] / world
                    rec["n1_value"] = n1["value"]
        except Exception:
            pass
    del ring, recs
    torch.cuda.empty_cache()
    torch.cuda.synchronize()

    # ---- the product's own sharding: ONE call from rank 0 over all N devices ----
    cpu_barrier()  # the other ranks wait on the CPU: nothing of theirs may sit on the GPUs rank 0 is about to use
    if rank == 0:
        rec["sharded_call"] = sharded_call(args, torch, pgb200, lib, dev, world, n, R, L, width, seed)
    cpu_barrier()
    return rec


def sharded_call(args, torch, pgb200, lib, dev, world, n, R, L, width, seed):
    with open("/proc/meminfo") as f:
        avail = int(dict(x.split(":") for x in f)["MemAvailable"].split()[0]) * 1024
    v_slice = args.config5_slice or 16384
    while v_slice > 512 and v_slice * L + (32768 + v_slice) * R > 0.6 * avail:
        v_slice //= 2
    # variants [v0, v0 + v_slice) straddle v = 34 360, where the reference's u32 record offset wraps (pfile.rs:165)
    v0 = max(0, 34360 - v_slice // 2) if not args.config5_variants else 0
    m_img = v0 + v_slice
    stream = torch.cuda.current_stream().cuda_stream
    t_setup = time.perf_counter()
    image = torch.empty(12 + m_img * R, dtype=torch.uint8, pin_memory=True)
    image[:12] = torch.from_numpy(np.frombuffer(synth.pgen_header(m_img, n), dtype=np.uint8).copy())
    d_img = torch.empty(v_slice * R + 64, dtype=torch.uint8, device=dev)
    step = min(8192, v_slice)
    for a in range(0, m_img, step):  # rows outside the slice are generated too: the image is a complete .pgen
        nb = min(step, m_img - a)
        pgb200._check(lib.pgb_dev_synth_records_fast(d_img.data_ptr(), R, seed, a, nb, n, stream), "synth")
        image[12 + a * R:12 + (a + nb) * R].copy_(d_img[:nb * R])
    pgb200._check(lib.pgb_dev_synth_records_fast(d_img.data_ptr(), R, seed, v0, v_slice, n, stream), "synth")
    d_img[v_slice * R:].zero_()
    total = v_slice * L
    h_out = torch.empty(total + 64, dtype=torch.uint8, pin_memory=True)
    var = np.arange(v0, v0 + v_slice, dtype=np.uint32)
    blob_np, off_np = synth.uniform_prefix_blob(v_slice, v0, width)
    setup_s = time.perf_counter() - t_setup
    devices = list(range(world))
    with pgb200.PgenFile(image_ptr=image.data_ptr(), image_bytes=image.numel()) as f:
        nb, st = f.export_gt_vcf_mem(var, None, blob_np, off_np, h_out.data_ptr(), total, devices=devices)  # warm-up: buffers
        t0 = time.perf_counter()
        nb, st = f.export_gt_vcf_mem(var, None, blob_np, off_np, h_out.data_ptr(), total, devices=devices)
        dt = time.perf_counter() - t0
    assert nb == total
    # bit-exact against the device-resident path on device 0, block by block
    torch.cuda.set_device(dev)
    d_blob = torch.from_numpy(blob_np).to(dev)
    d_off = torch.from_numpy(off_np.view(np.int64)).to(dev)
    blk = 1024
    d_ref = torch.empty(blk * L + 1024, dtype=torch.uint8, device=dev)
    d_cmp = torch.empty(blk * L, dtype=torch.uint8, device=dev)
    d_meta = torch.zeros((blk + 1) * 4, dtype=torch.int64, device=dev)
    d_scr = torch.zeros(lib.pgb_dev_index_scratch_bytes(blk) // 8 + 1, dtype=torch.int64, device=dev)
    ok = True
    for a in range(0, v_slice, blk):
        nbk = min(blk, v_slice - a)
        pgb200._check(lib.pgb_dev_index_lines(None, d_off.data_ptr() + 8 * a, 0, nbk, n, R, d_meta.data_ptr(),
                                              d_scr.data_ptr(), stream), "K1")
        pgb200._check(lib.pgb_dev_format_lines_ex(d_img.data_ptr() + a * R, R, d_meta.data_ptr(), nbk, d_blob.data_ptr(), 0, 0,
                                                  None, n, width, d_ref.data_ptr(), 0, stream), "K2")
        d_cmp[:nbk * L].copy_(h_out[a * L:(a + nbk) * L], non_blocking=True)
        ok = ok and bool(torch.equal(d_cmp[:nbk * L], d_ref[:nbk * L]))
    lib.pgb_release_buffers()
    return {"n_devices": world, "bit_exact": ok, "ms": dt * 1e3, "vcf_gb_per_s": total / dt / 1e9,
            "genotypes_per_s": v_slice * n / dt, "slice_variants": [int(v0), int(v0 + v_slice)], "slice_bytes": int(total),
            "record_offsets_cross_4GiB": bool((v0 + v_slice) * R > 2 ** 32 > v0 * R), "chunks": int(st.n_chunks),
            "api": "pgb_export_gt_vcf_mem(devices=[0..%d]) from one process: page-locked .pgen image in, page-locked body out" % (world - 1),
            "checked_against": "K1+K2 device-resident on device 0, every byte", "setup_s": setup_s}



def device_resident_line(args, torch, pgb200, lib, dev, rank, barrier, max_over_ranks, name):
    """Device-resident K1 + K2 of another workload (same method as the main one: W warm-up steps, then `steps` timed
    steps, CUDA events on the launching stream, max over ranks): the roofline of the kernels the default workload
    does not exercise (the shared-memory batch kernel on the gather-heavy and short-line shapes)."""
    wl = WORKLOADS[name]
    n, m, width, seed = wl["n"], wl["m"], wl["width"], wl["seed"]
    R = synth.record_size(n)
    row0 = rank * m
    stream = torch.cuda.current_stream().cuda_stream
    recs = torch.zeros(m * R + 64, dtype=torch.uint8, device=dev)
    pgb200._check(lib.pgb_dev_synth_records_fast(recs.data_ptr(), R, seed, row0, m, n, stream), "synth")
    if wl["mk"] is None:
        var, n_lines = None, m
        blob_np, off_np = synth.uniform_prefix_blob(m, row0, width)
    else:
        var = synth.subset_indices(42, m, wl["mk"])
        n_lines = len(var)
        b, _ = synth.uniform_prefix_blob(m, row0, width)
        blob_np = np.ascontiguousarray(b.reshape(m, width)[var]).reshape(-1)
        off_np = np.arange(n_lines + 1, dtype=np.uint64) * np.uint64(width)
    sam = None if wl["k"] is None else synth.subset_indices(41, n, wl["k"])
    K = n if sam is None else len(sam)
    total = int(off_np[-1]) + n_lines * (4 * K + 1)
    d_blob = torch.from_numpy(blob_np).to(dev)
    d_off = torch.from_numpy(off_np.view(np.int64)).to(dev)
    d_rows = None if var is None else torch.from_numpy(var.view(np.int32)).to(dev)
    d_kidx = None
    if sam is not None:
        mask = np.zeros(4 * R, np.uint8)
        mask[sam] = 1
        d_mask = torch.from_numpy(mask).to(dev)
        d_kidx = torch.zeros(4 * R + 8, dtype=torch.int32, device=dev)
        d_cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        pgb200._check(lib.pgb_dev_compact_samples(d_mask.data_ptr(), 4 * R, d_kidx.data_ptr(), d_cnt.data_ptr(), stream), "K0")
        assert int(d_cnt.item()) == K
    d_meta = torch.zeros((n_lines + 1) * 4, dtype=torch.int64, device=dev)
    d_scr = torch.zeros(lib.pgb_dev_index_scratch_bytes(n_lines) // 8 + 1, dtype=torch.int64, device=dev)
    d_out = torch.empty(total + 1024, dtype=torch.uint8, device=dev)
    variant = int(os.environ.get("PGB_K2_VARIANT", "0"), 0)

    def k1():
        pgb200._check(lib.pgb_dev_index_lines(None if d_rows is None else d_rows.data_ptr(), d_off.data_ptr(), 0, n_lines,
                                              K, R, d_meta.data_ptr(), d_scr.data_ptr(), stream), "K1")

    def k2():
        pgb200._check(lib.pgb_dev_format_lines_ex(recs.data_ptr(), R, d_meta.data_ptr(), n_lines, d_blob.data_ptr(), 0, 0,
                                                  None if d_kidx is None else d_kidx.data_ptr(), K, width, d_out.data_ptr(),
                                                  variant, stream), "K2")

    for _ in range(args.warmup):
        k1(); k2()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev0.record()
    for a, b in evs:
        k1()
        a.record()
        k2()
        b.record()
    ev1.record()
    barrier()
    dev_ms = max_over_ranks(ev0.elapsed_time(ev1))
    k2_ms = statistics.mean(a.elapsed_time(b) for a, b in evs)
    alg = n_lines * ((R if sam is None else min(R, 32 * len(np.unique(sam // 128)))) + width + (width + 4 * K + 1))
    peak, _ = peaks()
    # the first and the last line against what the kernel's inputs say (prefix bytes and the newline)
    L = width + 4 * K + 1
    ok = bytes(d_out[:width].cpu().numpy()) == bytes(blob_np[:width]) and int(d_out[total - 1].item()) == 10 and \
        bytes(d_out[total - L:total - L + width].cpu().numpy()) == bytes(blob_np[-width:])
    return {"workload": name, "description": wl["desc"], "kept_samples": K, "kept_variants_per_gpu": n_lines,
            "value": n_lines * K * args.steps / (dev_ms * 1e-3), "unit": UNIT + " per GPU", "ms_per_step": dev_ms / args.steps,
            "kernel": "k2_batch_kernel" if (sam is not None or width + 4 * K + 1 <= 3072) else "k2_format_kernel",
            "kernel_ms": k2_ms, "algorithmic_bytes_per_launch": int(alg), "achieved_gbs": alg / (k2_ms * 1e-3) / 1e9,
            "frac_of_measured_hbm_peak": alg / (k2_ms * 1e-3) / 1e9 / peak, "first_last_line_ok": ok}

# ------------------------------------------------------------------------ GPU arm ---
def run_b200(args, wl, rank, world, local_rank):
    import torch
    import pgb200  # fails loudly when libpgb200.so is missing: there is no fallback path

    if not torch.cuda.is_available() or pgb200.lib.pgb_device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device: libpgb200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)
        cpu_pg = dist.new_group(backend="gloo")  # CPU-side barriers: no kernel sits on a GPU while a rank waits
    lib = pgb200.lib
    stream = torch.cuda.current_stream().cuda_stream

    n, m, width, seed = wl["n"], wl["m"], wl["width"], wl["seed"]
    R = synth.record_size(n)
    row0 = rank * m  # this rank's contiguous variant range of the (world*m)-variant matrix
    # ---- inputs, resident in HBM ----
    recs = torch.zeros(m * R + 64, dtype=torch.uint8, device=dev)
    pgb200._check(lib.pgb_dev_synth_records(recs.data_ptr(), R, seed, row0, m, n, stream), "synth")
    if wl["mk"] is None:
        var = None
        n_lines = m
        blob_np, off_np = synth.uniform_prefix_blob(m, row0, width)
    else:
        var = synth.subset_indices(42, m, wl["mk"])
        n_lines = len(var)
        b, _ = synth.uniform_prefix_blob(m, row0, width)
        blob_np = np.ascontiguousarray(b.reshape(m, width)[var]).reshape(-1)
        off_np = np.arange(n_lines + 1, dtype=np.uint64) * np.uint64(width)
    sam = None if wl["k"] is None else synth.subset_indices(41, n, wl["k"])
    K = n if sam is None else len(sam)
    total = int(off_np[-1]) + n_lines * (4 * K + 1)
    d_blob = torch.from_numpy(blob_np).to(dev)
    d_off = torch.from_numpy(off_np.view(np.int64)).to(dev)
    d_rows = None if var is None else torch.from_numpy(var.view(np.int32)).to(dev)
    d_kidx = None
    if sam is not None:
        # K0 on the device: keep-mask -> kept-sample index list
        mask = np.zeros(4 * R, np.uint8)
        mask[sam] = 1
        d_mask = torch.from_numpy(mask).to(dev)
        d_kidx = torch.zeros(4 * R + 8, dtype=torch.int32, device=dev)
        d_cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        pgb200._check(lib.pgb_dev_compact_samples(d_mask.data_ptr(), 4 * R, d_kidx.data_ptr(), d_cnt.data_ptr(), stream), "K0")
        assert int(d_cnt.item()) == K
    d_meta = torch.zeros((n_lines + 1) * 4, dtype=torch.int64, device=dev)
    d_scr = torch.zeros(lib.pgb_dev_index_scratch_bytes(n_lines) // 8 + 1, dtype=torch.int64, device=dev)
    d_out = torch.empty(total + 1024, dtype=torch.uint8, device=dev)
    variant = int(os.environ.get("PGB_K2_VARIANT", "0"), 0)

    def k1():
        pgb200._check(lib.pgb_dev_index_lines(None if d_rows is None else d_rows.data_ptr(), d_off.data_ptr(), 0, n_lines,
                                              K, R, d_meta.data_ptr(), d_scr.data_ptr(), stream), "K1")

    def k2():
        pgb200._check(lib.pgb_dev_format_lines_ex(recs.data_ptr(), R, d_meta.data_ptr(), n_lines, d_blob.data_ptr(), 0, 0,
                                                  None if d_kidx is None else d_kidx.data_ptr(), K, width, d_out.data_ptr(),
                                                  variant, stream), "K2")

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def cpu_barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier(group=cpu_pg)

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def min_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    sampler = ClockSampler(local_rank).start() if rank == 0 else None

    # ---- device-resident: W warm-up steps, then exactly K timed steps ----
    for _ in range(args.warmup):
        k1(); k2()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev0.record()
    for a, b in evs:
        k1()
        a.record()
        k2()
        b.record()
    ev1.record()
    barrier()
    dev_ms = max_over_ranks(ev0.elapsed_time(ev1))
    k2_ms = statistics.mean(a.elapsed_time(b) for a, b in evs)
    launches_per_step = 4
    genotypes_step = n_lines * K
    value = world * genotypes_step * args.steps / (dev_ms * 1e-3)

    # ---- roofline of the dominant kernel (K2) ----
    # algorithmic bytes per kept variant: R (record) + P (prefix read) + L (line written)
    alg_bytes = n_lines * (R + width + (width + 4 * K + 1)) if sam is None else \
        n_lines * (min(R, 32 * len(np.unique(sam // 128))) + width + (width + 4 * K + 1))
    peak, peak_src = peaks()
    achieved = alg_bytes / (k2_ms * 1e-3) / 1e9
    # store-only ceiling (16-byte stores of a constant over the same output buffer)
    fill_bytes = total // 16 * 16
    for _ in range(2):
        lib.pgb_dev_fill(d_out.data_ptr(), fill_bytes, 0, stream)
    fa, fb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fa.record()
    for _ in range(5):
        lib.pgb_dev_fill(d_out.data_ptr(), fill_bytes, 0, stream)
    fb.record()
    torch.cuda.synchronize()
    fill_gbs = 5 * fill_bytes / (fa.elapsed_time(fb) * 1e-3) / 1e9
    traffic = traffic_src = None
    tp = os.path.join(ROOT, "profiles", "k2_traffic.json")
    if os.path.exists(tp):
        try:
            with open(tp) as f:
                tj = json.load(f)
            traffic = tj.get(args.workload)
            traffic_src = tj.get("source", "ncu, static (profiles/k2_traffic.json)") if traffic is not None else None
        except Exception:
            traffic = None

    # ---- end to end through the C ABI with host buffers ----
    if args.no_e2e:
        if rank == 0:
            print(json.dumps({"workload": args.workload, "value": value, "ms_per_step": dev_ms / args.steps, "k2_ms": k2_ms,
                              "achieved_gbs": achieved, "frac": achieved / peak, "store_only_ceiling_gbs": fill_gbs,
                              "variant": variant, "note": "kernel-only development line (--no-e2e); not a bench line"}), flush=True)
        if sampler:
            sampler.stop()
        if dist is not None:
            dist.destroy_process_group()
        return
    image = torch.empty(12 + m * R, dtype=torch.uint8, pin_memory=True)
    image[:12] = torch.from_numpy(np.frombuffer(synth.pgen_header(m, n), dtype=np.uint8).copy())
    image[12:].copy_(recs[:m * R])
    torch.cuda.synchronize()
    h_out = torch.empty(total + 64, dtype=torch.uint8, pin_memory=True)
    # host-side ceilings, measured in this run on all ranks at once (the host side is shared)
    cal = d2h_calibration(torch, d_out, h_out, total, barrier)
    d2h_gbs = cal["peak"]
    # aggregate ceiling of the box: all ranks' bytes over the time of the SLOWEST rank (summing per-rank rates would
    # credit the ranks that finish early with bandwidth the others were not yet using)
    agg_d2h = max(world * total / max_over_ranks(cal["whole_body_s"]), world * total / max_over_ranks(cal["chunked_s"])) / 1e9
    n_thr = max(1, min(16, (os.cpu_count() or 2) // max(1, world)))
    barrier()
    host_fill_gbs = host_write_probe(h_out.numpy(), n_thr)
    barrier()
    e2e = None
    with pgb200.PgenFile(image_ptr=image.data_ptr(), image_bytes=image.numel()) as f:
        def e2e_step():
            nb, st = f.export_gt_vcf_mem(var, sam, blob_np, off_np, h_out.data_ptr(), total, devices=[local_rank])
            assert nb == total
            return st
        for _ in range(args.warmup):
            st = e2e_step()
        barrier()
        t0 = time.perf_counter()
        e_launch = 0
        for _ in range(args.steps):
            st = e2e_step()
            e_launch += st.kernel_launches
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        barrier()
        e2e = {"value": world * genotypes_step * args.steps / e2e_s, "unit": UNIT,
               "h2d_bytes_per_step": int(st.bytes_h2d), "d2h_bytes_per_step": int(st.bytes_d2h),
               "ms_per_step": 1e3 * e2e_s / args.steps, "vcf_gb_per_s": world * total * args.steps / e2e_s / 1e9,
               "api": "pgb_export_gt_vcf_mem (page-locked .pgen image in, page-locked VCF body out)",
               "launches_per_step": int(e_launch // args.steps), "chunks_per_step": int(st.n_chunks),
               "roofline": {"bound": "pcie_d2h", "achieved": total * args.steps / e2e_s / 1e9, "peak": d2h_gbs,
                            "unit": "GB/s per GPU", "frac": total * args.steps / e2e_s / 1e9 / d2h_gbs,
                            "peak_source": "max of {one cudaMemcpyAsync of the whole body, the export's own pattern without kernels "
                                           "(128 MiB chunks over 3 streams)} device->pinned host, all ranks concurrently, measured in this run",
                            "whole_body_gbs": cal["whole_body_gbs"], "chunked_3_streams_128MiB_gbs": cal["chunked_3_streams_128MiB_gbs"],
                            "aggregate_peak_gbs_all_gpus": agg_d2h,
                            "aggregate_peak_source": "all ranks' bytes / slowest rank's time, max over the two copy patterns",
                            "aggregate_achieved_gbs_all_gpus": world * total * args.steps / e2e_s / 1e9,
                            "frac_of_aggregate_peak": world * total * args.steps / e2e_s / 1e9 / agg_d2h,
                            "host_memory_fill_gbs": host_fill_gbs, "host_memory_fill_threads": n_thr,
                            "host_memory_fill_note": "CPU threads of this rank filling the same page-locked buffer, all ranks at once"}}
        # parity spot check of the e2e result against the device-resident one (first/last 1 MiB)
        torch.cuda.synchronize()
        k1(); k2()
        torch.cuda.synchronize()
        for sl in (slice(0, 1 << 20), slice(max(0, total - (1 << 20)), total)):
            assert torch.equal(h_out[sl], d_out[sl].cpu()), "e2e and device-resident outputs differ"
        # ---- end to end INCLUDING the file write (north_star's second e2e form): the same call
        #      with an fd sink, as Pfile::output_vcf uses it.  Rank 0, N=1, 1 warm-up + 2 timed.
        # ---- end to end INCLUDING the file write (north_star's second e2e form): the same call with an fd sink,
        #      as Pfile::output_vcf uses it; every rank writes its own file at the same time.  1 warm-up + 2 timed.
        def file_leg(where, kind):
            ok = os.path.isdir(where)
            if ok:
                vfs = os.statvfs(where)
                ok = vfs.f_bavail * vfs.f_frsize >= world * total + (4 << 30)
            if min_over_ranks(1.0 if ok else 0.0) < 1.0:
                return None
            path = os.path.join(where, "pgb_bench_%d_%d.vcf" % (os.getpid(), rank))
            try:
                # raw sink rates for the same bytes, all ranks at once, on the same file system
                mv = memoryview(h_out.numpy())[:min(total, (4 << 30) // world)]
                barrier()
                ceil = file_sink_ceilings(path + ".probe", mv, n_thr)
                best = max([v for v in ceil.values() if isinstance(v, float)] or [0.0])
                agg_best = sum_over_ranks(best)
                od_rate = ceil.get("o_direct_pwrite_%d" % n_thr)
                agg_od = sum_over_ranks(od_rate if isinstance(od_rate, float) else 0.0)
                ts = []
                for it in range(3):
                    fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
                    try:
                        barrier()
                        t0 = time.perf_counter()
                        stf = f.export_gt_vcf(var, sam, blob_np, off_np, fd, devices=[local_rank])
                        ts.append(max_over_ranks(time.perf_counter() - t0))
                    finally:
                        os.close(fd)
                    assert os.path.getsize(path) == total
                with open(path, "rb") as fh:  # spot check against the in-memory result
                    assert fh.read(1 << 20) == bytes(h_out[:1 << 20].numpy())
                tt = sum(ts[1:]) / 2
                return {"value": world * genotypes_step / tt, "unit": UNIT, "vcf_gb_per_s": world * total / tt / 1e9,
                        "ms_per_step": 1e3 * tt, "sink": "one regular file per rank in %s (%s), no fsync" % (where, kind),
                        "api": "pgb_export_gt_vcf (fd sink)", "chunks": int(stf.n_chunks),
                        "roofline": {"bound": "storage", "achieved": world * total / tt / 1e9, "peak": agg_best, "unit": "GB/s (all ranks)",
                                     "frac": world * total / tt / 1e9 / agg_best if agg_best else None,
                                     "peak_source": "best raw sink rate of the same bytes without the GPU, all ranks at once, measured in this "
                                                    "run: max over {1 buffered pwrite thread, N buffered pwrite threads, N O_DIRECT pwrite "
                                                    "threads, N copies through a MAP_SHARED mapping}, N = %d per rank" % n_thr,
                                     "rank0_sink_rates_gbs": ceil, "o_direct_ceiling_gbs_all_ranks": agg_od}}
            finally:
                for q in (path, path + ".probe"):
                    if os.path.exists(q):
                        os.unlink(q)

        e2e_file = e2e_file_disk = None
        if not args.no_file:
            shm = os.environ.get("PGB_BENCH_FILE_DIR") or "/dev/shm"
            e2e_file = file_leg(shm, "tmpfs" if shm == "/dev/shm" else "PGB_BENCH_FILE_DIR")
            disk = tempfile.gettempdir()
            if os.path.isdir(disk) and os.stat(disk).st_dev != os.stat(shm).st_dev:
                e2e_file_disk = file_leg(disk, "block-device file system, page cache")
                if e2e_file_disk is not None:
                    os.environ["PGB_ODIRECT"] = "1"  # the opt-in O_DIRECT output stage on the same file system
                    try:
                        od = file_leg(disk, "block-device file system, O_DIRECT output stage (PGB_ODIRECT=1)")
                    finally:
                        del os.environ["PGB_ODIRECT"]
                    e2e_file_disk["o_direct_variant"] = None if od is None else dict(
                        {k: od[k] for k in ("value", "vcf_gb_per_s", "ms_per_step", "sink")},
                        parallel_o_direct_ceiling_gbs=od["roofline"]["o_direct_ceiling_gbs_all_ranks"],
                        frac_of_parallel_o_direct_ceiling=(od["vcf_gb_per_s"] / od["roofline"]["o_direct_ceiling_gbs_all_ranks"]
                                                           if od["roofline"]["o_direct_ceiling_gbs_all_ranks"] else None))
    clocks = sampler.stop() if sampler else None

    # ---- the other BASELINE shapes, device-resident (the kernels the default workload does not exercise) ----
    other = None
    if args.workload == "chr22" and not args.no_other:
        other = [device_resident_line(args, torch, pgb200, lib, dev, rank, barrier, max_over_ranks, w)
                 for w in ("gather", "random1", "biobank-block")]
        torch.cuda.empty_cache()

    # ---- configs[4] (biobank shape), strong scaling + the product's own multi-device call ----
    config5 = None
    if not args.no_config5:
        del d_out, recs, h_out, image, d_blob, d_off, d_meta, d_scr
        torch.cuda.empty_cache()
        config5 = run_config5(args, torch, pgb200, lib, dev, rank, world, barrier, cpu_barrier, max_over_ranks)

    # ---- CPU baseline beside it (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        with tempfile.TemporaryDirectory(prefix="pgb_cpu_") as td:
            ref = CpuReference(wl, td)
            rows = ref.rows_for_seconds(12.0)
            dt, g = ref.run(rows)
        cpu = {"value": g / dt, "unit": UNIT, "cores": 1, "kind": "port", "host_cores": os.cpu_count(),
               "sample": f"first {rows} of {m} variants ({g} genotypes, {dt:.1f} s), oracle/pgen_oracle.c reference-faithful I/O, output to a tmp file"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": args.workload, "description": wl["desc"], "n_samples": n, "variants_per_gpu": m,
                       "kept_samples": K, "kept_variants_per_gpu": n_lines, "prefix_bytes": width,
                       "vcf_body_bytes_per_gpu_step": total, "sharding": f"contiguous variant ranges x{world}, no collective",
                       "l2": "inputs (0.69 GB records) and output (>= 0.5 GB) exceed the 126 MB L2; no flush needed"},
            "vcf_gb_per_s": world * total * args.steps / (dev_ms * 1e-3) / 1e9,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": "k2_batch_kernel" if (sam is not None or width + 4 * K + 1 <= 3072) else "k2_format_kernel",
                         "kernel_ms": k2_ms,
                         "algorithmic_bytes_per_launch": int(alg_bytes), "peak_source": peak_src,
                         "frac_of_nominal_8TBs": achieved / 8000.0, "store_only_ceiling_gbs": fill_gbs},
            "e2e": e2e, "e2e_file": e2e_file, "e2e_file_disk": e2e_file_disk, "other_workloads": other, "config5": config5, "cpu_baseline": cpu, "gpu_launches": launches_per_step * args.steps, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="chr22", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-file", action="store_true", help="skip the e2e leg that writes a file")
    ap.add_argument("--n-samples", type=int, default=0, help="kernel development: override the workload's sample count")
    ap.add_argument("--n-variants", type=int, default=0, help="kernel development: override the workload's variant count")
    ap.add_argument("--no-e2e", action="store_true", help="kernel development: device-resident part only, short line")
    ap.add_argument("--no-config5", action="store_true", help="skip the configs[4] strong-scaling leg")
    ap.add_argument("--no-other", action="store_true", help="skip the device-resident lines of the other BASELINE shapes")
    ap.add_argument("--config5-variants", type=int, default=0, help="development: shrink configs[4] to this many variants")
    ap.add_argument("--config5-slice", type=int, default=0, help="variants of the slice given to the multi-device call (default 16384)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1 and args.impl == "b200":
        # plain `python bench.py --gpus N`: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29531"), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    wl = dict(WORKLOADS[args.workload])
    if args.n_samples:
        wl["n"] = args.n_samples
    if args.n_variants:
        wl["m"] = args.n_variants
    if args.impl == "reference":
        run_reference(args, wl, rank, world)
    else:
        run_b200(args, wl, rank, world, local_rank)


if __name__ == "__main__":
    main()
