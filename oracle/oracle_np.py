"""oracle_np.py — second, independent CPU restatement (numpy) of pgen-rs's export path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg as the checker.  The product (pgen-rs_b200/) never imports it.

PARITY UNPINNED BY THE REFERENCE (no goldens, reference is Rust and cannot be built
here); see oracle/pgen_oracle.c for how the oracle is pinned instead.  This file is
deliberately written in a different style from the C oracle (whole-matrix numpy
indexing instead of per-genotype loops) so that the two can check each other.

Reference citations are to /root/reference/src/pfile.rs unless stated.
"""
from __future__ import annotations

import io
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

MAGIC = b"\x6c\x1b"

# pfile.rs:177-183 — code -> text; pfile.rs:186-187 — '\t' + text = 4 bytes per genotype
_GT_FIELD = np.array(
    [list(b"\t0/0"), list(b"\t0/1"), list(b"\t1/1"), list(b"\t./.")], dtype=np.uint8
)


class OracleError(Exception):
    """Stands for the reference's panic (process exit 101)."""


# ------------------------------------------------------------------ .pgen ---

def read_pgen_header(pgen_path: str) -> Tuple[int, int]:
    """Pfile::from_prefix, pfile.rs:38-76 → (num_variants, num_samples)."""
    with open(pgen_path, "rb") as f:
        h = f.read(12)
    if len(h) < 2:
        raise OracleError("short header")
    if h[:2] != MAGIC:
        raise OracleError("bad magic (pfile.rs:47)")
    if len(h) < 3 or h[2] != 0x02:
        raise OracleError("storage mode != 0x02 (pfile.rs:53)")
    if len(h) < 12:
        raise OracleError("short header")
    m = int.from_bytes(h[3:7], "little")
    n = int.from_bytes(h[7:11], "little")
    if h[11] != 0x40:
        raise OracleError("flags != 0x40 (pfile.rs:69)")
    return m, n


def record_size(num_samples: int) -> int:
    """variant_record_size, pfile.rs:196-200 (u32 arithmetic)."""
    bit_size = (num_samples * 2) & 0xFFFFFFFF
    return (bit_size // 8) + (0 if bit_size % 8 == 0 else 1)


def record_offset(var_idx: int, rec_size: int, faithful_u32: bool = False) -> int:
    """pfile.rs:165; faithful_u32 reproduces the release-mode wrap."""
    if faithful_u32:
        return 12 + ((var_idx * rec_size) & 0xFFFFFFFF)
    return 12 + var_idx * rec_size


def decode_matrix(records: np.ndarray, sam_idx: np.ndarray) -> np.ndarray:
    """pfile.rs:172-175 for a (rows, R) u8 matrix and a sample index vector → (rows, K) codes."""
    sam_idx = np.asarray(sam_idx, dtype=np.int64)
    by = records[:, sam_idx >> 2]
    sh = ((sam_idx & 3) * 2).astype(np.uint8)
    return (by >> sh[None, :]) & 3


def gt_text_block(codes: np.ndarray) -> np.ndarray:
    """(rows, K) codes → (rows, 4K) bytes of '\\t' + GT (pfile.rs:177-188)."""
    rows, k = codes.shape
    return _GT_FIELD[codes].reshape(rows, 4 * k)


# ------------------------------------------------------- .pvar / .psam ------

def _read_line_spans(data: bytes):
    """BufRead::read_line spans: up to and including b'\\n'."""
    pos = 0
    n = len(data)
    while True:
        j = data.find(b"\n", pos)
        e = n if j < 0 else j + 1
        yield pos, e
        if e >= n:
            # subsequent reads return empty strings forever
            while True:
                yield n, n
        pos = e


def read_pvar_header(data: bytes) -> Tuple[bytes, bytes]:
    """read_pvar_header, pfile.rs:202-220 → (comment lines verbatim, column-name line)."""
    lines = []
    for b, e in _read_line_spans(data):
        if e > b and data[b:b + 1] == b"#":
            lines.append(data[b:e])
        else:
            break
    if not lines:
        raise OracleError("no '#' header line (pfile.rs:217 unwrap)")
    col = lines.pop()
    return b"".join(lines), col


def metadata_header_start(data: bytes) -> int:
    """find_metadata_file_header_start, pfile.rs:248-268."""
    prev = 0
    cur = 0
    for b, e in _read_line_spans(data):
        prev, cur = cur, e - b
        if not (cur > 0 and data[b:b + 1] == b"#"):
            offset = cur + prev - 1
            if offset < 0:
                raise OracleError("u64 underflow in header seek")
            return e - offset
    raise AssertionError


def parse_table(data: bytes) -> Tuple[List[str], List[List[str]]]:
    """metadata_file_reader, pfile.rs:270-283: csv (tab, has_headers, non-flexible) on the
    bytes after the '#' of the column line.  Quoting is not restated: a '\"' is rejected."""
    start = min(metadata_header_start(data), len(data))
    body = data[start:]
    if b'"' in body:
        raise OracleError("input needs csv quoting rules; rejected")
    text = body.replace(b"\r\n", b"\n").replace(b"\r", b"\n")
    recs = [ln for ln in text.split(b"\n") if ln != b""]
    if not recs:
        return [], []
    try:
        rows = [[f.decode("utf-8") for f in r.split(b"\t")] for r in recs]
    except UnicodeDecodeError as ex:  # StringRecord requires UTF-8
        raise OracleError("invalid utf-8") from ex
    ncol = len(rows[0])
    for r in rows:
        if len(r) != ncol:
            raise OracleError("ragged row (csv UnequalLengths)")
    return rows[0], rows[1:]


# -------------------------------------------------- evalexpr 11.3.0 subset ---
# Call sites: pfile.rs:87-98 (query), :322-328 (filter).  Every column is bound as a
# String.  Restated: literals (string with \" and \\ escapes, int, float, true/false),
# identifiers, unary ! and -, binary ^ * / % + - < > <= >= == != && ||, parentheses,
# with evalexpr's precedences (120 ^, 110 unary, 100 */%, 95 +-, 80 comparisons,
# 75 &&, 70 ||).  Operands are evaluated eagerly (no short circuit), as evalexpr does.

class ExprError(OracleError):
    pass


_SPECIAL = set('+-*/%^()=!<>&|,;"')


def _parse_literal(word: str):
    """evalexpr's literal rule: decimal or 0x-hex i64, else f64, else bool, else identifier."""
    try:
        if word.startswith("0x"):
            v = int(word[2:], 16)
        elif word.isascii() and word.isdigit():
            v = int(word)
        else:
            raise ValueError
        if v >= 2 ** 63:
            raise ValueError
        return ("int", v)
    except ValueError:
        pass
    if any(ch.isdigit() for ch in word) and all(ch in "0123456789.eE+-" for ch in word):
        try:
            return ("float", float(word))
        except ValueError:
            pass
    if word == "true":
        return ("bool", True)
    if word == "false":
        return ("bool", False)
    return ("id", word)


def _tokenize(src: str):
    i, n = 0, len(src)
    out = []
    while i < n:
        c = src[i]
        if c in " \t\n\r\x0b\x0c":
            i += 1
        elif c == '"':
            i += 1
            buf = []
            while True:
                if i >= n:
                    raise ExprError("unterminated string")
                c = src[i]
                if c == "\\":
                    if i + 1 < n and src[i + 1] in '"\\':
                        buf.append(src[i + 1])
                        i += 2
                    else:
                        raise ExprError("illegal escape sequence")
                elif c == '"':
                    i += 1
                    break
                else:
                    buf.append(c)
                    i += 1
            out.append(("str", "".join(buf)))
        elif src.startswith(("==", "!=", "<=", ">=", "&&", "||"), i):
            out.append(("op", src[i:i + 2]))
            i += 2
        elif c in "+-*/%^<>!()":
            out.append(("op", c))
            i += 1
        elif c in _SPECIAL:
            raise ExprError(f"unsupported operator {c!r}")
        else:
            j = i
            while j < n and src[j] not in _SPECIAL and src[j] not in " \t\n\r\x0b\x0c":
                j += 1
            out.append(_parse_literal(src[i:j]))
            i = j
    return out


_BIN_PREC = {
    "^": 120, "*": 100, "/": 100, "%": 100, "+": 95, "-": 95,
    "<": 80, ">": 80, "<=": 80, ">=": 80, "==": 80, "!=": 80, "&&": 75, "||": 70,
}


class _Parser:
    def __init__(self, toks):
        self.t = toks
        self.i = 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else None

    def parse(self):
        if not self.t:
            raise ExprError("empty expression")
        node = self.expr(0)
        if self.peek() is not None:
            raise ExprError("trailing tokens")
        return node

    def expr(self, min_prec):
        lhs = self.unary()
        while True:
            tk = self.peek()
            if tk is None or tk[0] != "op" or tk[1] not in _BIN_PREC:
                return lhs
            prec = _BIN_PREC[tk[1]]
            if prec < min_prec:
                return lhs
            self.i += 1
            # '^' is right-associative in evalexpr, everything else left
            rhs = self.expr(prec if tk[1] == "^" else prec + 1)
            lhs = ("bin", tk[1], lhs, rhs)

    def unary(self):
        tk = self.peek()
        if tk is None:
            raise ExprError("unexpected end")
        if tk[0] == "op" and tk[1] in "!-":
            self.i += 1
            return ("un", tk[1], self.expr(110))
        if tk[0] == "op" and tk[1] == "(":
            self.i += 1
            node = self.expr(0)
            nx = self.peek()
            if nx is None or nx != ("op", ")"):
                raise ExprError("unmatched parenthesis")
            self.i += 1
            return node
        if tk[0] in ("str", "int", "float", "bool"):
            self.i += 1
            return ("lit", tk[1])
        if tk[0] == "id":
            self.i += 1
            nx = self.peek()
            if nx is not None and (nx[0] in ("str", "int", "float", "bool", "id") or nx == ("op", "(")):
                raise ExprError("function calls are not part of the restated subset")
            return ("var", tk[1])
        raise ExprError(f"unexpected token {tk!r}")


def compile_expr(src: str):
    return _Parser(_tokenize(src)).parse()


def _is_num(v):
    return isinstance(v, (int, float)) and not isinstance(v, bool)


def _i64(v):
    """evalexpr integers are i64 with checked arithmetic."""
    if isinstance(v, int) and not isinstance(v, bool) and not (-2 ** 63 <= v < 2 ** 63):
        raise ExprError("integer overflow")
    return v


def eval_node(node, ctx):
    kind = node[0]
    if kind == "lit":
        return node[1]
    if kind == "var":
        if node[1] not in ctx:
            raise ExprError(f"variable {node[1]!r} not found")
        return ctx[node[1]]
    if kind == "un":
        v = eval_node(node[2], ctx)
        if node[1] == "!":
            if not isinstance(v, bool):
                raise ExprError("expected boolean")
            return not v
        if not _is_num(v):
            raise ExprError("expected number")
        return _i64(-v)
    op, a, b = node[1], eval_node(node[2], ctx), eval_node(node[3], ctx)
    if op in ("&&", "||"):
        if not isinstance(a, bool) or not isinstance(b, bool):
            raise ExprError("expected boolean")
        return (a and b) if op == "&&" else (a or b)
    if op in ("==", "!="):
        same = type(a) is type(b) and a == b
        return same if op == "==" else not same
    if op in ("<", ">", "<=", ">="):
        if isinstance(a, str) and isinstance(b, str):
            pass
        elif _is_num(a) and _is_num(b):
            pass
        else:
            raise ExprError("expected two numbers or two strings")
        return {"<": a < b, ">": a > b, "<=": a <= b, ">=": a >= b}[op]
    if op == "+":
        if isinstance(a, str) and isinstance(b, str):
            return a + b
        if _is_num(a) and _is_num(b):
            return _i64(a + b)
        raise ExprError("expected two numbers or two strings")
    if not (_is_num(a) and _is_num(b)):
        raise ExprError("expected number")
    if op == "-":
        return _i64(a - b)
    if op == "*":
        return _i64(a * b)
    if op == "/":
        if isinstance(a, int) and isinstance(b, int):
            if b == 0:
                raise ExprError("division by zero")
            q = abs(a) // abs(b)
            return q if (a >= 0) == (b >= 0) else -q
        return a / b
    if op == "%":
        if isinstance(a, int) and isinstance(b, int):
            if b == 0:
                raise ExprError("modulo by zero")
            r = abs(a) % abs(b)
            return r if a >= 0 else -r
        import math
        return math.fmod(a, b)
    if op == "^":
        return float(a) ** float(b)
    raise ExprError(op)


def filter_metadata(headers: Sequence[str], rows: Sequence[Sequence[str]], query: Optional[str]) -> List[int]:
    """filter_metadata, pfile.rs:312-335 → ascending list of kept row indices."""
    if query is None:
        return list(range(len(rows)))
    tree = compile_expr(query)
    kept = []
    for idx, r in enumerate(rows):
        ctx = dict(zip(headers, r))
        v = eval_node(tree, ctx)
        if not isinstance(v, bool):
            raise ExprError("expected boolean")
        if v:
            kept.append(idx)
    return kept


def query_metadata(headers, rows, query: Optional[str], fstring: str) -> List[str]:
    """query_metadata, pfile.rs:78-102 → the lines println! would print."""
    qt = compile_expr(query) if query is not None else None
    ft = compile_expr(fstring)
    out = []
    for r in rows:
        ctx = dict(zip(headers, r))
        ok = True
        if qt is not None:
            ok = eval_node(qt, ctx)
            if not isinstance(ok, bool):
                raise ExprError("expected boolean")
        if ok:
            s = eval_node(ft, ctx)
            if not isinstance(s, str):
                raise ExprError("expected string")
            out.append(s)
    return out


# ------------------------------------------------------------- output_vcf ---

_RUST_ASCII_WS = b" \t\n\x0b\x0c\r"


def vcf_header(pvar_data: bytes, iids: Sequence[str]) -> bytes:
    """pfile.rs:139-146."""
    comments, col = read_pvar_header(pvar_data)
    return (
        b"##fileformat=VCFv4.2\n##source=pgen-rs\n"
        + comments
        + col.strip(_RUST_ASCII_WS)
        + b"\tFORMAT\t"
        + "\t".join(iids).encode()
        + b"\n"
    )


def line_prefix(row: Sequence[str]) -> bytes:
    """pfile.rs:157-161: every field followed by a tab, then 'GT'."""
    return b"".join(f.encode() + b"\t" for f in row) + b"GT"


def format_body(records: np.ndarray, var_rows: Sequence[int], sam_idx: Sequence[int],
                prefixes: Sequence[bytes], chunk: int = 4096) -> bytes:
    """pfile.rs:156-192 on an in-memory (M, R) record matrix."""
    out = io.BytesIO()
    sam_idx = np.asarray(sam_idx, dtype=np.int64)
    var_rows = np.asarray(var_rows, dtype=np.int64)
    for c0 in range(0, len(var_rows), chunk):
        vr = var_rows[c0:c0 + chunk]
        if len(sam_idx):
            txt = gt_text_block(decode_matrix(records[vr], sam_idx))
        else:
            txt = np.zeros((len(vr), 0), dtype=np.uint8)
        for k in range(len(vr)):
            out.write(prefixes[c0 + k])
            out.write(txt[k].tobytes())
            out.write(b"\n")
    return out.getvalue()


def load_records(pgen_path: str, faithful_u32: bool = False) -> Tuple[np.ndarray, int, int]:
    m, n = read_pgen_header(pgen_path)
    r = record_size(n)
    raw = np.fromfile(pgen_path, dtype=np.uint8, offset=12)
    if faithful_u32:
        raise NotImplementedError("use records_at for the wrapped-offset mode")
    if raw.size < m * r:
        # the reference never checks the length; missing rows only fail when touched
        pad = np.zeros(m * r - raw.size, dtype=np.uint8)
        raw = np.concatenate([raw, pad])
    return raw[: m * r].reshape(m, r), m, n


def read_pgen10(pgen_path: str):
    """Standard-format (.pgen storage mode 0x10) header walk: what Pgen::from_file_path computes
    (src/pgen.rs:18-137: header-format byte :50-67, block count :100-102, block-offset table
    :104-110, main-header-body size :116-133, records offset :135-137) plus the per-variant
    record index its check_main_header_body walks past (:172-258), with that walker's three
    defects corrected (last block = M - 65536*b, not M % 65536; lengths are little-endian
    integers of record_length_bytes bytes; allele-count bytes are part of the body).
    Returns (M, N, off[M+1], typ[M], length[M]).  Test infrastructure only."""
    data = open(pgen_path, "rb").read()
    if data[0:2] != b"\x6c\x1b":
        raise OracleError("bad magic")
    if data[2] != 0x10:
        raise OracleError("not storage mode 0x10")
    m = int.from_bytes(data[3:7], "little")
    n = int.from_bytes(data[7:11], "little")
    fmt = data[11]
    mode = fmt & 0xF
    if mode // 4 > 1:
        raise OracleError("unsupported record storage mode")
    type_bits = 4 if mode // 4 == 0 else 8
    len_bytes = mode % 4 + 1
    ac_bytes = (fmt >> 4) & 3
    if (fmt >> 6) & 3 != 1:
        raise OracleError("provisional_ref_storage != 1 (pgen.rs:66)")
    nb = (m + 65535) // 65536
    block_off = [int.from_bytes(data[12 + 8 * b:20 + 8 * b], "little") for b in range(nb)]
    pos = 12 + 8 * nb
    off = np.zeros(m + 1, dtype=np.uint64)
    typ = np.zeros(m, dtype=np.uint8)
    length = np.zeros(m, dtype=np.uint32)
    v = 0
    for b in range(nb):
        cnt = 65536 if b + 1 < nb else m - 65536 * b
        tb = (cnt * type_bits + 7) // 8
        tbytes = data[pos:pos + tb]
        lbytes = data[pos + tb:pos + tb + cnt * len_bytes]
        pos += tb + cnt * len_bytes + cnt * ac_bytes
        o = block_off[b]
        for k in range(cnt):
            t = tbytes[k] if type_bits == 8 else (tbytes[k // 2] >> ((k & 1) * 4)) & 0xF
            ln = int.from_bytes(lbytes[k * len_bytes:(k + 1) * len_bytes], "little")
            off[v], typ[v], length[v] = o, t, ln
            o += ln
            v += 1
        if b + 1 == nb:
            off[m] = o
    if nb == 0:
        off[0] = pos
    return m, n, off, typ, length


def records_standard(pgen_path: str, rows: Sequence[int]) -> np.ndarray:
    """The plain 2-bit records (type 0, ceil(2N/8) bytes) of the given variants of a mode-0x10 file."""
    m, n, off, typ, length = read_pgen10(pgen_path)
    r = record_size(n)
    data = np.fromfile(pgen_path, dtype=np.uint8)
    out = np.zeros((len(rows), r), dtype=np.uint8)
    for i, v in enumerate(rows):
        if typ[v] != 0 or length[v] != r:
            raise OracleError("variant %d is not a plain 2-bit record" % v)
        out[i] = data[int(off[v]):int(off[v]) + r]
    return out


def output_vcf(prefix: str, sam_query: Optional[str], var_query: Optional[str], out_path: str,
               var_idx: Optional[Sequence[int]] = None, sam_idx: Optional[Sequence[int]] = None) -> None:
    """Pfile::output_vcf, pfile.rs:104-194.  Selections come from the queries unless explicit
    index lists are given (then the queries must be None)."""
    m, n = read_pgen_header(prefix + ".pgen")
    r = record_size(n)
    pvar_data = open(prefix + ".pvar", "rb").read()
    psam_data = open(prefix + ".psam", "rb").read()
    read_pvar_header(pvar_data)  # pfile.rs:110 panics first if there is no header
    sh, srows = parse_table(psam_data)
    if "IID" not in sh:
        raise OracleError("IID not among the headers (pfile.rs:125)")
    iid_col = sh.index("IID")
    vh, vrows = parse_table(pvar_data)
    vi = list(var_idx) if var_idx is not None else filter_metadata(vh, vrows, var_query)
    si = list(sam_idx) if sam_idx is not None else filter_metadata(sh, srows, sam_query)
    for s in si:
        if s // 4 >= r:
            raise OracleError("sample index out of record (pfile.rs:173)")
    size = os.path.getsize(prefix + ".pgen")
    with open(out_path, "wb") as out:
        out.write(vcf_header(pvar_data, [srows[s][iid_col] for s in si]))
        if not vi:
            return
        raw = np.memmap(prefix + ".pgen", dtype=np.uint8, mode="r")
        sidx = np.asarray(si, dtype=np.int64)
        for c0 in range(0, len(vi), 2048):
            rows = vi[c0:c0 + 2048]
            recs = np.empty((len(rows), r), dtype=np.uint8)
            for k, v in enumerate(rows):
                off = record_offset(v, r)
                if off + r > size:
                    raise OracleError("read_exact past EOF (pfile.rs:170)")
                recs[k] = raw[off:off + r]
            pre = [line_prefix(vrows[v]) for v in rows]
            out.write(format_body(recs, range(len(rows)), sidx, pre))
