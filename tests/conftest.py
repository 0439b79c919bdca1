"""Shared test plumbing.

Layout of the suite
  -m "not gpu": the oracle against the hand-derived golden vectors and against its numpy
                twin; the C++ host mirror (selection, header, prefixes, expressions, CLI
                argument surface); a lane-by-lane host simulation of K2's byte logic; that
                libpgb200.so loads and exports every symbol include/pgb200.h declares; the
                world_size-2 gloo test of the variant-range sharding used by bench.py.
  -m gpu      : parity proper — the CUDA path through the C ABI against the oracle.

Nothing here (or in any test) reads /root/reference at run time.
"""
import ctypes
import gzip
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools"), os.path.join(ROOT, "pgen-rs_b200", "python"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _run(cmd, **kw):
    r = subprocess.run(cmd, capture_output=True, text=True, **kw)
    if r.returncode != 0:
        raise RuntimeError(f"{' '.join(cmd)} failed:\n{r.stdout}\n{r.stderr}")


@pytest.fixture(scope="session")
def orc():
    """ctypes handle of the C oracle (oracle/pgen_oracle.c), built on demand."""
    so = os.path.join(ROOT, "oracle", "_build", "liborc.so")
    src = os.path.join(ROOT, "oracle", "pgen_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        _run(["make", "-C", os.path.join(ROOT, "oracle")])
    lib = ctypes.CDLL(so)
    lib.orc_output_vcf.restype = ctypes.c_int
    lib.orc_output_vcf.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                   ctypes.c_char_p, ctypes.c_int]
    lib.orc_export_body.restype = ctypes.c_int
    lib.orc_export_body.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    lib.orc_read_pgen_header.restype = ctypes.c_int
    lib.orc_read_pgen_header.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32)]
    lib.orc_record_size.restype = ctypes.c_uint32
    lib.orc_record_size.argtypes = [ctypes.c_uint32]
    lib.orc_record_offset.restype = ctypes.c_uint64
    lib.orc_record_offset.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int]
    lib.orc_format_gt_fields.restype = ctypes.c_uint64
    lib.orc_format_gt_fields.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p]
    lib.orc_strerror.restype = ctypes.c_char_p
    return lib


def orc_output_vcf(orc, prefix, var_idx, sam_idx, out_path, flags=2):
    """flags: 1 = faithful u32 offsets, 2 = bulk I/O (default for tests)."""
    vi = None if var_idx is None else np.ascontiguousarray(var_idx, dtype=np.int64)
    si = None if sam_idx is None else np.ascontiguousarray(sam_idx, dtype=np.int64)
    dummy = np.zeros(1, dtype=np.int64)
    rc = orc.orc_output_vcf(
        prefix.encode(),
        None if vi is None else (vi.ctypes.data if len(vi) else dummy.ctypes.data), -1 if vi is None else len(vi),
        None if si is None else (si.ctypes.data if len(si) else dummy.ctypes.data), -1 if si is None else len(si),
        out_path.encode(), flags)
    return rc


@pytest.fixture(scope="session")
def k2sim():
    """TEST-ONLY host simulation of K2 (tests/hostsim/k2_hostsim.cpp)."""
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "k2sim.so")
    srcs = [os.path.join(ROOT, "tests", "hostsim", "k2_hostsim.cpp"),
            os.path.join(ROOT, "pgen-rs_b200", "csrc", "k2_core.cuh"),
            os.path.join(ROOT, "pgen-rs_b200", "csrc", "k2_batch.cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        _run(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-shared", "-fPIC", "-o", so, srcs[0]])
    lib = ctypes.CDLL(so)
    lib.sim_format_lines.restype = ctypes.c_int
    lib.sim_format_lines.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int]
    lib.sim_format_lines_sfx.restype = ctypes.c_int
    lib.sim_format_lines_sfx.argtypes = lib.sim_format_lines.argtypes + [ctypes.c_uint32, ctypes.c_uint32]
    lib.sim_format_lines_batch.restype = ctypes.c_int
    lib.sim_format_lines_batch.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint64,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p,
                                           ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int]
    return lib


@pytest.fixture(scope="session")
def pgb():
    """The product binding; importing it fails loudly when libpgb200.so is missing."""
    import pgb200
    return pgb200


@pytest.fixture(scope="session")
def kat_cases():
    with open(os.path.join(ROOT, "tests", "golden", "kat.json")) as f:
        return json.load(f)["cases"]


def write_case(case, dirpath):
    prefix = os.path.join(str(dirpath), case["name"])
    with open(prefix + ".pgen", "wb") as f:
        f.write(bytes.fromhex(case["pgen_hex"].replace(" ", "")))
    with open(prefix + ".pvar", "wb") as f:
        f.write(case["pvar"].encode())
    with open(prefix + ".psam", "wb") as f:
        f.write(case["psam"].encode())
    return prefix


@pytest.fixture(scope="session")
def basic1(tmp_path_factory):
    """data/basic1 of the reference: its real .pvar/.psam (shipped gzip'd under tests/data)
    plus a synthesised .pgen (the reference's own blob is absent from its checkout), seed 1."""
    import synth
    d = tmp_path_factory.mktemp("basic1")
    prefix = os.path.join(str(d), "basic1")
    data = os.path.join(ROOT, "tests", "data")
    with open(prefix + ".pvar", "wb") as f:
        f.write(gzip.open(os.path.join(data, "basic1.pvar.gz")).read())
    with open(prefix + ".psam", "wb") as f:
        f.write(gzip.open(os.path.join(data, "basic1.psam.txt.gz")).read())
    synth.write_pgen(prefix + ".pgen", 1, 17784, 2504)
    return prefix


@pytest.fixture(scope="session")
def random1(tmp_path_factory):
    """data/random1 shape: real .psam (300 samples), synthetic 5-column .pvar and .pgen, seed 2."""
    import synth
    d = tmp_path_factory.mktemp("random1")
    prefix = os.path.join(str(d), "random1")
    data = os.path.join(ROOT, "tests", "data")
    with open(prefix + ".psam", "wb") as f:
        f.write(gzip.open(os.path.join(data, "random1.psam.txt.gz")).read())
    synth.write_pvar(prefix + ".pvar", "random1", 200000, 2)
    synth.write_pgen(prefix + ".pgen", 2, 200000, 300)
    return prefix


def has_gpu():
    try:
        import pgb200
        return pgb200.lib.pgb_device_count() > 0
    except Exception:
        return False
