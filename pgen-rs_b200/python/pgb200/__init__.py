"""ctypes binding of libpgb200.so (include/pgb200.h) for tests/ and bench.py.

This is plumbing only: every function here forwards to the C ABI.  There is no Python or
CPU implementation of the hot path — if the shared library is missing the import fails
loudly (run `python -c "import __graft_entry__ as g; g.build()"` or `make -C pgen-rs_b200`).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.abspath(os.path.join(_HERE, "..", ".."))
# PGB_LIB: an alternative build of the same library (A/B comparisons of kernel variants on one box)
LIB_PATH = os.environ.get("PGB_LIB") or os.path.join(PKG_ROOT, "lib", "libpgb200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: libpgb200 has no fallback path; build it with `make -C {PKG_ROOT}`"
    )
lib = C.CDLL(LIB_PATH)


class PgbError(RuntimeError):
    def __init__(self, status: int, where: str):
        self.status = status
        detail = lib.pgb_last_error().decode(errors="replace")
        msg = lib.pgb_strerror(status).decode()
        super().__init__(f"{where}: {msg} ({status})" + (f": {detail}" if detail else ""))


# status codes (include/pgb200.h)
OK, E_IO, E_MAGIC, E_MODE, E_FLAGS, E_ARG, E_RANGE, E_NO_DEVICE, E_CUDA, E_NOMEM = 0, -1, -2, -3, -4, -5, -6, -7, -8, -9
E_NO_HEADER, E_NO_IID, E_CSV, E_EXPR, E_SPACE = -10, -11, -12, -13, -14


class Stats(C.Structure):
    _fields_ = [
        ("n_lines", C.c_uint64), ("n_kept_samples", C.c_uint64), ("genotypes", C.c_uint64),
        ("bytes_out", C.c_uint64), ("bytes_h2d", C.c_uint64), ("bytes_d2h", C.c_uint64),
        ("kernel_launches", C.c_uint64), ("device_ms", C.c_double), ("e2e_ms", C.c_double),
        ("n_devices", C.c_int32), ("n_chunks", C.c_int32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class LineMeta(C.Structure):
    _fields_ = [("line_off", C.c_uint64), ("rec_off", C.c_uint64), ("pfx_off", C.c_uint64),
                ("pfx_len", C.c_uint32), ("reserved", C.c_uint32)]


class Pgen10Info(C.Structure):
    _fields_ = [
        ("n_variants", C.c_uint32), ("n_samples", C.c_uint32), ("storage_mode", C.c_uint8),
        ("header_format", C.c_uint8), ("record_type_bits", C.c_uint8), ("record_length_bytes", C.c_uint8),
        ("allele_count_bytes", C.c_uint8), ("provisional_ref_storage", C.c_uint8),
        ("variant_block_count", C.c_uint32), ("variant_block_offsets_offset", C.c_uint64),
        ("main_header_body_offset", C.c_uint64), ("main_header_body_size", C.c_uint64),
        ("variant_records_offset", C.c_uint64),
    ]


# every symbol include/pgb200.h declares: (restype, argtypes)
_vp, _u64, _u32, _i = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
SYMBOLS = {
    "pgb_open": (_i, [C.c_char_p, C.POINTER(_vp)]),
    "pgb_open_mem": (_i, [_vp, _u64, C.POINTER(_vp)]),
    "pgb_open_standard": (_i, [C.c_char_p, C.POINTER(_vp)]),
    "pgb_dims": (None, [_vp, C.POINTER(_u32), C.POINTER(_u32), C.POINTER(_u32)]),
    "pgb_close": (None, [_vp]),
    "pgb_record_bytes": (_u32, [_u32]),
    "pgb_record_offset": (_u64, [_u64, _u32]),
    "pgb_export_gt_vcf": (_i, [_vp, _vp, _u64, _vp, _u64, _vp, _vp, _i, _vp, _i, C.POINTER(Stats)]),
    "pgb_export_gt_vcf_mem": (_i, [_vp, _vp, _u64, _vp, _u64, _vp, _vp, _vp, _u64, C.POINTER(_u64), _vp, _i,
                                   C.POINTER(Stats)]),
    "pgb_export_gt_vcf_rows": (_i, [_vp, _vp, _u64, _vp, _u64, _vp, _u64, _vp, _vp, _i, _vp, _i, C.POINTER(Stats)]),
    "pgb_export_gt_vcf_rows_mem": (_i, [_vp, _vp, _u64, _vp, _u64, _vp, _u64, _vp, _vp, _vp, _u64, C.POINTER(_u64), _vp, _i,
                                        C.POINTER(Stats)]),
    "pgb_release_buffers": (None, []),
    "pgb_body_bytes": (_u64, [_u64, _u64, _vp]),
    "pgb_shard_plan": (_i, [_u64, _u64, _vp, _i, _vp, _vp]),
    "pgb_pfile_output_vcf": (_i, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, _vp, _i, C.POINTER(Stats)]),
    "pgb_pfile_query": (_i, [C.c_char_p, C.c_char_p, C.c_char_p, _i, _i]),
    "pgb_plan_vcf": (_i, [C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(_vp)]),
    "pgb_plan_free": (None, [_vp]),
    "pgb_plan_n_var": (_u64, [_vp]),
    "pgb_plan_n_sam": (_u64, [_vp]),
    "pgb_plan_var_idx": (_vp, [_vp]),
    "pgb_plan_sam_idx": (_vp, [_vp]),
    "pgb_plan_header": (_vp, [_vp, C.POINTER(_u64)]),
    "pgb_plan_pvar_text": (_vp, [_vp, C.POINTER(_u64)]),
    "pgb_plan_row_off": (_vp, [_vp]),
    "pgb_plan_row_len": (_vp, [_vp]),
    "pgb_plan_prefix_blob": (_vp, [_vp, C.POINTER(_u64)]),
    "pgb_plan_prefix_off": (_vp, [_vp]),
    "pgb_pgen10_index": (_i, [C.c_char_p, C.POINTER(Pgen10Info), _vp, _vp, _vp]),
    "pgb_dev_compact_samples": (_i, [_vp, _u32, _vp, _vp, _vp]),
    "pgb_dev_index_scratch_bytes": (_u64, [_u64]),
    "pgb_dev_index_lines": (_i, [_vp, _vp, _u64, _u64, _u32, _u64, _vp, _vp, _vp]),
    "pgb_dev_index_lines_off": (_i, [_vp, _vp, _u64, _u64, _u32, _vp, _vp, _vp]),
    "pgb_dev_format_lines": (_i, [_vp, _vp, _u64, _vp, _vp, _u32, _u32, _vp, _i, _vp]),
    "pgb_dev_index_lines_ex": (_i, [_vp, _vp, _u64, _vp, _vp, _u32, _u64, _u64, _u32, _vp, _vp, _vp]),
    "pgb_dev_format_lines_ex": (_i, [_vp, _u32, _vp, _u64, _vp, _u32, _u32, _vp, _u32, _u32, _vp, _i, _vp]),
    "pgb_dev_synth_records": (_i, [_vp, _u64, _u64, _u64, _u64, _u32, _vp]),
    "pgb_dev_synth_records_fast": (_i, [_vp, _u64, _u32, _u64, _u64, _u32, _vp]),
    "pgb_dev_fill": (_i, [_vp, _u64, _i, _vp]),
    "pgb_dev_fill_pattern": (_i, [_vp, _u64, _u32, _u32, _u32, _u32, _i, _vp]),
    "pgb_device_count": (_i, []),
    "pgb_strerror": (C.c_char_p, [_i]),
    "pgb_last_error": (C.c_char_p, []),
    "pgb_abi_version": (_i, []),
}
for _name, (_res, _args) in SYMBOLS.items():
    _fn = getattr(lib, _name)  # AttributeError here = the library does not export the symbol
    _fn.restype = _res
    _fn.argtypes = _args


def _check(rc: int, where: str) -> None:
    if rc != OK:
        raise PgbError(rc, where)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


_EMPTY_U32 = np.zeros(1, dtype=np.uint32)


def _sam_ptr(si: Optional[np.ndarray]):
    # NULL means "all samples"; an empty selection must still pass a non-NULL pointer
    if si is None:
        return None
    return si.ctypes.data if len(si) else _EMPTY_U32.ctypes.data


def _u32arr(x) -> Optional[np.ndarray]:
    return None if x is None else np.ascontiguousarray(x, dtype=np.uint32)


class PgenFile:
    """Handle returned by pgb_open / pgb_open_mem."""

    def __init__(self, path: Optional[str] = None, image: Optional[np.ndarray] = None, image_ptr: int = 0,
                 image_bytes: int = 0, standard: bool = False):
        h = _vp()
        self._keep = None
        if path is not None and standard:
            _check(lib.pgb_open_standard(os.fsencode(path), C.byref(h)), f"pgb_open_standard({path})")
        elif path is not None:
            _check(lib.pgb_open(os.fsencode(path), C.byref(h)), f"pgb_open({path})")
        elif image is not None:
            self._keep = image
            _check(lib.pgb_open_mem(image.ctypes.data, image.nbytes, C.byref(h)), "pgb_open_mem")
        else:
            _check(lib.pgb_open_mem(image_ptr, image_bytes, C.byref(h)), "pgb_open_mem")
        self.h = h
        m, n, r = _u32(), _u32(), _u32()
        lib.pgb_dims(h, C.byref(m), C.byref(n), C.byref(r))
        self.n_variants, self.n_samples, self.record_bytes = m.value, n.value, r.value

    def close(self):
        if self.h:
            lib.pgb_close(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def export_gt_vcf(self, var_idx, sam_idx, prefix_blob: np.ndarray, prefix_off: np.ndarray, out_fd: int,
                      devices: Optional[Sequence[int]] = None, n_var: Optional[int] = None) -> Stats:
        vi, si = _u32arr(var_idx), _u32arr(sam_idx)
        po = np.ascontiguousarray(prefix_off, dtype=np.uint64)
        pb = np.ascontiguousarray(prefix_blob, dtype=np.uint8)
        nv = len(po) - 1 if n_var is None else n_var
        dv = None if devices is None else np.ascontiguousarray(devices, dtype=np.int32)
        st = Stats()
        rc = lib.pgb_export_gt_vcf(self.h, _ptr(vi), nv, _sam_ptr(si),
                                   0 if si is None else len(si), _ptr(pb), _ptr(po), out_fd, _ptr(dv),
                                   0 if dv is None else len(dv), C.byref(st))
        _check(rc, "pgb_export_gt_vcf")
        return st

    def export_gt_vcf_mem(self, var_idx, sam_idx, prefix_blob, prefix_off, out_ptr: int, out_cap: int,
                          devices: Optional[Sequence[int]] = None, n_var: Optional[int] = None):
        vi, si = _u32arr(var_idx), _u32arr(sam_idx)
        po = np.ascontiguousarray(prefix_off, dtype=np.uint64)
        pb = np.ascontiguousarray(prefix_blob, dtype=np.uint8)
        nv = len(po) - 1 if n_var is None else n_var
        dv = None if devices is None else np.ascontiguousarray(devices, dtype=np.int32)
        st = Stats()
        n_out = _u64()
        rc = lib.pgb_export_gt_vcf_mem(self.h, _ptr(vi), nv, _sam_ptr(si),
                                       0 if si is None else len(si), _ptr(pb), _ptr(po), out_ptr, out_cap,
                                       C.byref(n_out), _ptr(dv), 0 if dv is None else len(dv), C.byref(st))
        _check(rc, "pgb_export_gt_vcf_mem")
        return n_out.value, st


def export_rows_to_bytes(f: PgenFile, var_idx, sam_idx, pvar_text: np.ndarray, row_off, row_len, devices=None) -> bytes:
    """pgb_export_gt_vcf_rows_mem: prefixes built on the device from the raw .pvar image."""
    vi, si = _u32arr(var_idx), _u32arr(sam_idx)
    ro = np.ascontiguousarray(row_off, dtype=np.uint64)
    rl = np.ascontiguousarray(row_len, dtype=np.uint32)
    tx = np.ascontiguousarray(pvar_text, dtype=np.uint8)
    k = f.n_samples if sam_idx is None else len(sam_idx)
    total = int(rl.astype(np.uint64).sum()) + len(rl) * (3 + 4 * k + 1)
    out = np.empty(max(total, 1), dtype=np.uint8)
    dv = None if devices is None else np.ascontiguousarray(devices, dtype=np.int32)
    st = Stats()
    n_out = _u64()
    rc = lib.pgb_export_gt_vcf_rows_mem(f.h, _ptr(vi), len(rl), _sam_ptr(si), 0 if si is None else len(si), _ptr(tx), tx.nbytes,
                                        _ptr(ro), _ptr(rl), out.ctypes.data, total, C.byref(n_out), _ptr(dv),
                                        0 if dv is None else len(dv), C.byref(st))
    _check(rc, "pgb_export_gt_vcf_rows_mem")
    return out[:n_out.value].tobytes()


def export_to_bytes(f: PgenFile, var_idx, sam_idx, prefix_blob, prefix_off, devices=None) -> bytes:
    po = np.ascontiguousarray(prefix_off, dtype=np.uint64)
    k = f.n_samples if sam_idx is None else len(sam_idx)
    total = lib.pgb_body_bytes(len(po) - 1, k, po.ctypes.data)
    out = np.empty(max(total, 1), dtype=np.uint8)
    n, _ = f.export_gt_vcf_mem(var_idx, sam_idx, prefix_blob, po, out.ctypes.data, total, devices)
    return out[:n].tobytes()


def shard_plan(n_kept_samples: int, prefix_off, n_shards: int):
    """pgb_shard_plan: (line_begin, byte_begin), each n_shards + 1 entries."""
    po = np.ascontiguousarray(prefix_off, dtype=np.uint64)
    lb = np.zeros(n_shards + 1, np.uint64)
    bb = np.zeros(n_shards + 1, np.uint64)
    _check(lib.pgb_shard_plan(len(po) - 1, n_kept_samples, po.ctypes.data, n_shards, lb.ctypes.data, bb.ctypes.data),
           "pgb_shard_plan")
    return lb, bb


class VcfPlan:
    """CPU planning stage of Pfile::output_vcf (pgb_plan_vcf)."""

    def __init__(self, prefix: str, sam_query: Optional[str], var_query: Optional[str]):
        h = _vp()
        enc = lambda s: None if s is None else s.encode()
        _check(lib.pgb_plan_vcf(os.fsencode(prefix), enc(sam_query), enc(var_query), C.byref(h)), "pgb_plan_vcf")
        try:
            nv, ns = lib.pgb_plan_n_var(h), lib.pgb_plan_n_sam(h)
            self.var_idx = np.ctypeslib.as_array(C.cast(lib.pgb_plan_var_idx(h), C.POINTER(_u32)), (nv,)).copy() if nv else np.zeros(0, np.uint32)
            self.sam_idx = np.ctypeslib.as_array(C.cast(lib.pgb_plan_sam_idx(h), C.POINTER(_u32)), (ns,)).copy() if ns else np.zeros(0, np.uint32)
            ln = _u64()
            p = lib.pgb_plan_pvar_text(h, C.byref(ln))
            self.pvar_text = np.frombuffer(C.string_at(p, ln.value), dtype=np.uint8).copy() if ln.value else np.zeros(0, np.uint8)
            self.row_off = np.ctypeslib.as_array(C.cast(lib.pgb_plan_row_off(h), C.POINTER(_u64)), (nv,)).copy() if nv else np.zeros(0, np.uint64)
            self.row_len = np.ctypeslib.as_array(C.cast(lib.pgb_plan_row_len(h), C.POINTER(_u32)), (nv,)).copy() if nv else np.zeros(0, np.uint32)
            p = lib.pgb_plan_header(h, C.byref(ln))
            self.header = C.string_at(p, ln.value)
            p = lib.pgb_plan_prefix_blob(h, C.byref(ln))
            self.prefix_blob = np.frombuffer(C.string_at(p, ln.value), dtype=np.uint8).copy() if ln.value else np.zeros(0, np.uint8)
            self.prefix_off = np.ctypeslib.as_array(C.cast(lib.pgb_plan_prefix_off(h), C.POINTER(_u64)), (nv + 1,)).copy()
        finally:
            lib.pgb_plan_free(h)


def pfile_output_vcf(prefix: str, sam_query: Optional[str], var_query: Optional[str], out_path: str,
                     devices: Optional[Sequence[int]] = None) -> Stats:
    enc = lambda s: None if s is None else s.encode()
    dv = None if devices is None else np.ascontiguousarray(devices, dtype=np.int32)
    st = Stats()
    rc = lib.pgb_pfile_output_vcf(os.fsencode(prefix), enc(sam_query), enc(var_query), os.fsencode(out_path), _ptr(dv),
                                  0 if dv is None else len(dv), C.byref(st))
    _check(rc, "pgb_pfile_output_vcf")
    return st


def pfile_query(prefix: str, fstring: str, query: Optional[str], samples: bool, out_fd: int) -> None:
    rc = lib.pgb_pfile_query(os.fsencode(prefix), fstring.encode(), None if query is None else query.encode(),
                             1 if samples else 0, out_fd)
    _check(rc, "pgb_pfile_query")


def pgen10_index(path: str, want_index: bool = True):
    info = Pgen10Info()
    _check(lib.pgb_pgen10_index(os.fsencode(path), C.byref(info), None, None, None), "pgb_pgen10_index")
    if not want_index:
        return info, None, None, None
    m = info.n_variants
    off = np.zeros(m + 1, np.uint64)
    typ = np.zeros(max(m, 1), np.uint8)
    ln = np.zeros(max(m, 1), np.uint32)
    _check(lib.pgb_pgen10_index(os.fsencode(path), C.byref(info), off.ctypes.data, typ.ctypes.data, ln.ctypes.data),
           "pgb_pgen10_index")
    return info, off, typ[:m], ln[:m]


# ---- device-resident helpers (torch tensors supply device memory and streams) ----

def dev_format(records, pitch: int, var_row, prefix_blob, prefix_off, kidx, n_kept: int, max_prefix_len: int, out,
               meta, scratch, variant: int = 0, stream: int = 0, prefix_base: int = 0, n_lines: Optional[int] = None,
               record_bytes: Optional[int] = None):
    """K1 + K2 on torch CUDA tensors (raw data_ptr()s are passed through the C ABI)."""
    n = (prefix_off.numel() - 1) if n_lines is None else n_lines
    rc = lib.pgb_dev_index_lines(None if var_row is None else var_row.data_ptr(), prefix_off.data_ptr(), prefix_base, n,
                                 n_kept, pitch, meta.data_ptr(), scratch.data_ptr(), stream)
    _check(rc, "pgb_dev_index_lines")
    rc = lib.pgb_dev_format_lines_ex(records.data_ptr(), pitch if record_bytes is None else record_bytes, meta.data_ptr(), n,
                                     prefix_blob.data_ptr(), 0, 0, None if kidx is None else kidx.data_ptr(), n_kept,
                                     max_prefix_len, out.data_ptr(), variant, stream)
    _check(rc, "pgb_dev_format_lines_ex")
