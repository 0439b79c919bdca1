"""Seeded synthetic pfile generators for the BASELINE.json configs.

The genotype generator is a counter-based integer hash, so the same bytes can be
produced here (numpy) and on the device (pgb_dev_synth_records in
pgen-rs_b200/csrc/synth.cu) without moving data over PCIe; tests check the two agree.

Distribution: per-variant ALT frequency p ~ U(0.01, 0.5), codes drawn
Hardy-Weinberg (1-p)^2 / 2p(1-p) / p^2 with 1 % missing (code 3).  Integer-only:

    hv   = splitmix64(seed * 0x2545F4914F6CDD1D + v)                 (v = file row index)
    p16  = 655 + (((hv >> 40) * 32113) >> 24)                        (p in 1/65536 units)
    q16  = 65536 - p16
    t0   = (q16*q16*64881) >> 32 ;  t1 = t0 + ((2*p16*q16*64881) >> 32)
    hb   = splitmix64(hv ^ ((j+1) * 0xD6E8FEB86659FD93))             (j = byte in record)
    u_k  = (hb >> 16k) & 0xFFFF, k = 0..3                            (sample 4j+k)
    code = 3 if u_k >= 64881 else (u_k >= t0) + (u_k >= t1); padding samples (>= N) are 0
"""
from __future__ import annotations

import os
from typing import Optional, Sequence, Tuple

import numpy as np

U64 = np.uint64
_MIX_V = U64(0x2545F4914F6CDD1D)
_MIX_B = U64(0xD6E8FEB86659FD93)


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=U64)
    with np.errstate(over="ignore"):
        z = x + U64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> U64(30))) * U64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> U64(27))) * U64(0x94D049BB133111EB)
        return z ^ (z >> U64(31))


def record_size(n_samples: int) -> int:
    return (2 * n_samples + 7) // 8


def synth_records(seed: int, row0: int, n_rows: int, n_samples: int) -> np.ndarray:
    """(n_rows, R) uint8 records of file rows row0 .. row0+n_rows-1."""
    r = record_size(n_samples)
    out = np.empty((n_rows, r), dtype=np.uint8)
    if n_rows == 0 or r == 0:
        return out
    step = max(1, (1 << 22) // max(r, 1))
    jj = (np.arange(r, dtype=U64) + U64(1))
    with np.errstate(over="ignore"):
        jmix = jj * _MIX_B
        for a in range(0, n_rows, step):
            b = min(n_rows, a + step)
            v = np.arange(row0 + a, row0 + b, dtype=U64)
            hv = splitmix64(U64(seed) * _MIX_V + v)
            p16 = U64(655) + (((hv >> U64(40)) * U64(32113)) >> U64(24))
            q16 = U64(65536) - p16
            t0 = (q16 * q16 * U64(64881)) >> U64(32)
            t1 = t0 + ((U64(2) * p16 * q16 * U64(64881)) >> U64(32))
            hb = splitmix64(hv[:, None] ^ jmix[None, :])
            byte = np.zeros((b - a, r), dtype=np.uint8)
            for k in range(4):
                u = (hb >> U64(16 * k)) & U64(0xFFFF)
                code = (u >= t0[:, None]).astype(np.uint8) + (u >= t1[:, None]).astype(np.uint8)
                code[u >= U64(64881)] = 3
                byte |= (code << np.uint8(2 * k)).astype(np.uint8)
            out[a:b] = byte
    # padding bits of the last byte are zero
    rem = n_samples % 4
    if rem:
        out[:, r - 1] &= np.uint8((1 << (2 * rem)) - 1)
    return out


def _fmix32(h: np.ndarray) -> np.ndarray:
    h = h.astype(np.uint32)
    with np.errstate(over="ignore"):
        h ^= h >> np.uint32(16)
        h *= np.uint32(0x85EBCA6B)
        h ^= h >> np.uint32(13)
        h *= np.uint32(0xC2B2AE35)
        h ^= h >> np.uint32(16)
    return h


def synth_records_fast(seed: int, row0: int, n_rows: int, n_samples: int) -> np.ndarray:
    """The cheap generator of the biobank shape (pgb_dev_synth_records_fast is its device twin): word w (bytes
    4w..4w+3, little-endian) of file row v is fmix32(seed*0x9E3779B1 ^ v*0x85EBCA77 ^ (w+1)*0xC2B2AE3D) in 32-bit
    arithmetic (fmix32 = murmur3's finaliser); padding samples (>= N) are 0.  Uniform over the four codes."""
    r = record_size(n_samples)
    words = (r + 3) // 4
    out = np.empty((n_rows, r), dtype=np.uint8)
    if n_rows == 0 or r == 0:
        return out
    with np.errstate(over="ignore"):
        v = (np.arange(row0, row0 + n_rows, dtype=np.uint64) & U64(0xFFFFFFFF)).astype(np.uint32)
        w1 = np.arange(1, words + 1, dtype=np.uint32)
        s = np.uint32((seed * 0x9E3779B1) & 0xFFFFFFFF)
        h = _fmix32(s ^ (v * np.uint32(0x85EBCA77))[:, None] ^ (w1 * np.uint32(0xC2B2AE3D))[None, :])
    out[:] = h.astype("<u4").view(np.uint8).reshape(n_rows, 4 * words)[:, :r]
    rem = n_samples % 4
    if rem:
        out[:, r - 1] &= np.uint8((1 << (2 * rem)) - 1)
    return out


def pgen_header(n_variants: int, n_samples: int) -> bytes:
    """Appendix A.1: 6C 1B 02 | M le32 | N le32 | 40."""
    return b"\x6c\x1b\x02" + int(n_variants).to_bytes(4, "little") + int(n_samples).to_bytes(4, "little") + b"\x40"


def write_pgen(path: str, seed: int, n_variants: int, n_samples: int, block: int = 1 << 16) -> None:
    with open(path, "wb") as f:
        f.write(pgen_header(n_variants, n_samples))
        for a in range(0, n_variants, block):
            n = min(block, n_variants - a)
            f.write(synth_records(seed, a, n, n_samples).tobytes())


def write_pgen_bytes(path: str, records: np.ndarray, n_samples: int) -> None:
    with open(path, "wb") as f:
        f.write(pgen_header(records.shape[0], n_samples))
        f.write(np.ascontiguousarray(records, dtype=np.uint8).tobytes())


_ACGT = "ACGT"


def pvar_rows(kind: str, row0: int, n_rows: int, seed: int = 0):
    """Yield the tab-joined text of .pvar data rows (no newline).

    kind: 'random1' 5 columns (1 / i+1 / snp{i} / REF / ALT),
          'lean'    8 columns with '.' INFO (~40-byte prefix),
          '1000g'   8 columns with an AC/AF/AN-style INFO (~160-byte prefix).
    """
    idx = np.arange(row0, row0 + n_rows, dtype=U64)
    with np.errstate(over="ignore"):
        h = splitmix64(U64(seed) * _MIX_V + idx + U64(0x5EED))
    ref = (h & U64(3)).astype(np.int64)
    alt = ((ref + 1 + ((h >> U64(2)) % U64(3)).astype(np.int64)) % 4)
    hh = h.astype(object)
    for k in range(n_rows):
        i = row0 + k
        r, a = _ACGT[ref[k]], _ACGT[alt[k]]
        if kind == "random1":
            yield f"1\t{i + 1}\tsnp{i}\t{r}\t{a}"
        elif kind == "lean":
            yield f"22\t{16050000 + 32 * i + int(hh[k] >> 59)}\trs{i}\t{r}\t{a}\t.\tPASS\t."
        elif kind == "1000g":
            x = int(hh[k])
            ac = (x >> 8) % 5008
            info = (
                f"AC={ac};AF={ac / 5008:.6g};AN=5008;NS=2504;DP={(x >> 24) % 30000};"
                f"EAS_AF={((x >> 30) % 10000) / 10000:.4g};AMR_AF={((x >> 36) % 10000) / 10000:.4g};"
                f"AFR_AF={((x >> 42) % 10000) / 10000:.4g};EUR_AF={((x >> 48) % 10000) / 10000:.4g};"
                f"SAS_AF={((x >> 54) % 10000) / 10000:.4g};AA=.|||;VT=SNP"
            )
            yield f"22\t{16050000 + 32 * i + (x >> 59)}\trs{i}\t{r}\t{a}\t100\tPASS\t{info}"
        else:
            raise ValueError(kind)


_PVAR_COLS = {
    "random1": "#CHROM\tPOS\tID\tREF\tALT",
    "lean": "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO",
    "1000g": "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO",
}


def write_pvar(path: str, kind: str, n_variants: int, seed: int = 0, comments: Sequence[str] = ()) -> None:
    with open(path, "w", newline="") as f:
        for c in comments:
            f.write(c + "\n")
        f.write(_PVAR_COLS[kind] + "\n")
        for a in range(0, n_variants, 1 << 16):
            n = min(1 << 16, n_variants - a)
            f.write("\n".join(pvar_rows(kind, a, n, seed)))
            f.write("\n")


def write_psam(path: str, n_samples: int, cols: str = "#IID\tSEX") -> None:
    ncol = cols.count("\t") + 1
    with open(path, "w", newline="") as f:
        f.write(cols + "\n")
        for a in range(0, n_samples, 1 << 16):
            n = min(1 << 16, n_samples - a)
            f.write("\n".join("\t".join([f"per{i}"] + ["NA"] * (ncol - 1)) for i in range(a, a + n)))
            f.write("\n")


def prefix_blob(kind: str, rows: Sequence[int], seed: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """The bytes pfile.rs:157-161 writes before the genotype fields for the given file
    rows (each field + '\\t', then 'GT'), concatenated, plus the n+1 offsets — the
    prefix_blob / prefix_off arguments of pgb_export_gt_vcf."""
    rows = np.asarray(rows, dtype=np.int64)
    parts = []
    off = np.zeros(len(rows) + 1, dtype=np.uint64)
    pos = 0
    k = 0
    # contiguous runs are generated in bulk
    i = 0
    while i < len(rows):
        j = i
        while j + 1 < len(rows) and rows[j + 1] == rows[j] + 1:
            j += 1
        for txt in pvar_rows(kind, int(rows[i]), j - i + 1, seed):
            b = (txt + "\tGT").encode()
            parts.append(b)
            pos += len(b)
            k += 1
            off[k] = pos
        i = j + 1
    blob = np.frombuffer(b"".join(parts), dtype=np.uint8).copy() if parts else np.zeros(0, dtype=np.uint8)
    return blob, off


def uniform_prefix_blob(n: int, row0: int = 0, width: int = 40) -> Tuple[np.ndarray, np.ndarray]:
    """Fast vectorised fixed-width prefixes for the large bench shapes: an 8-column row
    `22 <POS> rs<id> A C . PASS .` + tab + `GT`, POS/id zero-padded so every prefix is
    exactly `width` bytes (>= 40)."""
    if width < 40:
        raise ValueError("width >= 40")
    # layout: "22\t" POS(9) "\t" "rs" ID(w) "\t" "A\tC\t.\tPASS\t.\t" "GT"
    fixed_tail = b"\tA\tC\t.\tPASS\t.\tGT"
    idw = width - (3 + 9 + 1 + 2 + len(fixed_tail))
    if idw < 1:
        raise ValueError("width too small")
    rows = np.arange(row0, row0 + n, dtype=np.int64)
    line = np.empty((n, width), dtype=np.uint8)
    line[:, 0:3] = np.frombuffer(b"22\t", dtype=np.uint8)
    pos = 16050000 + 32 * rows
    for d in range(9):
        line[:, 3 + 8 - d] = (pos % 10 + 48).astype(np.uint8)
        pos //= 10
    line[:, 12] = 9
    line[:, 13:15] = np.frombuffer(b"rs", dtype=np.uint8)
    ids = rows.copy()
    for d in range(idw):
        line[:, 15 + idw - 1 - d] = (ids % 10 + 48).astype(np.uint8)
        ids //= 10
    line[:, 15 + idw:] = np.frombuffer(fixed_tail, dtype=np.uint8)
    off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(width))
    return line.reshape(-1), off


def make_pfile(prefix: str, seed: int, n_variants: int, n_samples: int, pvar_kind: str = "lean",
               psam_cols: str = "#IID\tSEX", comments: Sequence[str] = ()) -> str:
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    write_pgen(prefix + ".pgen", seed, n_variants, n_samples)
    write_pvar(prefix + ".pvar", pvar_kind, n_variants, seed, comments)
    write_psam(prefix + ".psam", n_samples, psam_cols)
    return prefix


def subset_indices(seed: int, n: int, k: int) -> np.ndarray:
    """Exactly k of n, drawn without replacement with PCG64(seed), sorted ascending (cfg 4)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return np.sort(rng.choice(n, size=k, replace=False)).astype(np.uint32)


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser(description="write a synthetic pfile triple")
    ap.add_argument("prefix")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--variants", type=int, required=True)
    ap.add_argument("--samples", type=int, required=True)
    ap.add_argument("--pvar", default="lean", choices=list(_PVAR_COLS))
    a = ap.parse_args()
    make_pfile(a.prefix, a.seed, a.variants, a.samples, a.pvar)
