"""The C++ host mirror of the reference's Pfile API (pgen-rs_b200/host/) on the CPU: the
selections, VCF header and line prefixes it hands to the CUDA path must equal the oracle's,
its expression evaluator must agree with the oracle's independent restatement of the
evalexpr subset, and the CLI must mirror src/cli.rs.  No GPU compute is called here."""
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_np as onp
import synth
from conftest import ROOT, write_case

CLI = os.path.join(ROOT, "bin", "pgen-b200")


def test_library_exports_every_declared_symbol(pgb):
    hdr = open(os.path.join(ROOT, "include", "pgb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pgb_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"pgb_status"}
    assert declared, "no declarations parsed"
    assert declared == set(pgb.SYMBOLS), declared ^ set(pgb.SYMBOLS)
    for name in declared:
        getattr(pgb.lib, name)
    assert pgb.lib.pgb_abi_version() == 1
    assert pgb.lib.pgb_strerror(-7).decode().startswith("no usable CUDA device")


def test_geometry_matches_oracle(pgb):
    for n in [0, 1, 3, 4, 5, 300, 2504, 500000, 2**31 - 1]:
        assert pgb.lib.pgb_record_bytes(n) == onp.record_size(n)
    # u64 offsets (the reference's u32 product would wrap here, pfile.rs:165)
    assert pgb.lib.pgb_record_offset(34360, 125000) == 12 + 34360 * 125000 == onp.record_offset(34360, 125000)


def test_open_checks_in_reference_order(pgb, tmp_path):
    good = synth.pgen_header(3, 5) + bytes(6)
    cases = {b"": pgb.E_IO, good[:1]: pgb.E_IO, b"\x6c\x1c" + good[2:]: pgb.E_MAGIC,
             good[:2] + b"\x10" + good[3:]: pgb.E_MODE, good[:11] + b"\x00": pgb.E_FLAGS, good[:11]: pgb.E_IO}
    for i, (blob, rc) in enumerate(cases.items()):
        p = tmp_path / f"h{i}.pgen"
        p.write_bytes(blob)
        with pytest.raises(pgb.PgbError) as ei:
            pgb.PgenFile(str(p))
        assert ei.value.status == rc
        if len(blob) >= 12 or rc != pgb.E_IO:
            img = np.frombuffer(blob + bytes(16), dtype=np.uint8)
            with pytest.raises(pgb.PgbError) as ei:
                pgb.PgenFile(image=img[:max(len(blob), 1)])
            assert ei.value.status == rc
    with pytest.raises(pgb.PgbError) as ei:
        pgb.PgenFile(str(tmp_path / "missing.pgen"))
    assert ei.value.status == pgb.E_IO
    p = tmp_path / "ok.pgen"
    p.write_bytes(good)
    with pgb.PgenFile(str(p)) as f:
        assert (f.n_variants, f.n_samples, f.record_bytes) == (3, 5, 2)


def test_plan_matches_golden_vectors(pgb, kat_cases, tmp_path):
    for case in kat_cases:
        prefix = write_case(case, tmp_path)
        plan = pgb.VcfPlan(prefix, case["sam_query"], case["var_query"])
        assert plan.var_idx.tolist() == case["var_idx"], case["name"]
        assert plan.sam_idx.tolist() == case["sam_idx"], case["name"]
        vcf = case["vcf"].encode()
        assert vcf.startswith(plan.header), case["name"]
        # every line of the golden body starts with the planned prefix
        body = vcf[len(plan.header):]
        lines = body.split(b"\n")[:-1]
        assert len(lines) == len(plan.var_idx)
        for i, ln in enumerate(lines):
            pre = bytes(plan.prefix_blob[int(plan.prefix_off[i]):int(plan.prefix_off[i + 1])])
            assert ln.startswith(pre) and len(ln) == len(pre) + 4 * len(plan.sam_idx), case["name"]


def test_plan_basic1_config1(pgb, basic1):
    plan = pgb.VcfPlan(basic1, 'IID == "NA20900"', 'ALT == "G"')
    pvar = open(basic1 + ".pvar", "rb").read()
    vh, vrows = onp.parse_table(pvar)
    sh, srows = onp.parse_table(open(basic1 + ".psam", "rb").read())
    vi = onp.filter_metadata(vh, vrows, 'ALT == "G"')
    assert plan.var_idx.tolist() == vi and len(vi) == 4130
    assert plan.sam_idx.tolist() == [2444]
    assert plan.header == onp.vcf_header(pvar, ["NA20900"]) and len(plan.header) == 11866
    exp = b"".join(onp.line_prefix(vrows[v]) for v in vi)
    assert plan.prefix_blob.tobytes() == exp
    assert int(plan.prefix_off[-1]) == len(exp)
    assert pgb.lib.pgb_body_bytes(len(vi), 1, plan.prefix_off.ctypes.data) == 706753
    # what output_vcf hands to the device instead of the blob: the raw .pvar image and the kept rows in it
    assert plan.pvar_text.tobytes() == pvar
    for k in (0, 1, 77, len(vi) - 1):
        a, n = int(plan.row_off[k]), int(plan.row_len[k])
        assert pvar[a:a + n] + b"\tGT" == onp.line_prefix(vrows[vi[k]])


def test_plan_keep_all(pgb, basic1):
    plan = pgb.VcfPlan(basic1, None, None)
    assert plan.var_idx.tolist() == list(range(17784)) and plan.sam_idx.tolist() == list(range(2504))
    assert len(plan.header) + pgb.lib.pgb_body_bytes(17784, 2504, plan.prefix_off.ctypes.data) == 181130024


def test_plan_errors(pgb, tmp_path):
    prefix = str(tmp_path / "e")
    synth.write_pgen_bytes(prefix + ".pgen", np.zeros((2, 2), np.uint8), 5)
    open(prefix + ".pvar", "wb").write(b"#CHROM\tPOS\n1\t2\n1\t3\n")
    good_psam = b"#IID\tSEX\np0\t1\np1\t2\n"
    open(prefix + ".psam", "wb").write(good_psam)

    def status(sam=None, var=None):
        try:
            pgb.VcfPlan(prefix, sam, var)
            return 0
        except pgb.PgbError as e:
            return e.status
    assert status() == 0
    assert status(var='POS == ') == pgb.E_EXPR           # parse error
    assert status(var='POS') == pgb.E_EXPR               # not a boolean
    assert status(var='NOPE == "1"') == pgb.E_EXPR       # unknown variable
    assert status(var='POS > 1') == pgb.E_EXPR           # string vs int comparison errors in evalexpr
    assert status(sam='SEX == "1" && true') == 0
    open(prefix + ".psam", "wb").write(b"#FID\tSEX\nf\t1\n")
    assert status() == pgb.E_NO_IID
    open(prefix + ".psam", "wb").write(b"#IID\tSEX\np0\t1\np1\n")
    assert status() == pgb.E_CSV
    open(prefix + ".psam", "wb").write(b'#IID\tSEX\np0\t"1"\n')
    assert status() == pgb.E_CSV
    open(prefix + ".psam", "wb").write(good_psam)
    open(prefix + ".pvar", "wb").write(b"1\t2\n")
    assert status() == pgb.E_NO_HEADER
    os.remove(prefix + ".pvar")
    assert status() == pgb.E_IO


EXPRS = [
    'ALT == "G"', 'ALT != "G"', 'POS=="16647494" || POS=="51241285"', 'POS!="16647494" || POS!="51241285"',
    'REF == "A" && ALT == "C" || ID == "rs5"', 'REF == "A" && (ALT == "C" || ID == "rs5")', '!(REF == "A")',
    'CHROM + ":" + POS == "22:16050075"', 'ID < "rs3"', 'ID >= "rs3" && ID <= "rs7"', 'true', 'false || REF > "C"',
    '1 + 2 == 3 && REF == "T"', '2 ^ 3 == 8.0', '7 / 2 == 3 && 7 % 2 == 1 && -7 / 2 == -3', '1.5 * 2 == 3.0', '1 == 1.0',
    '"a\\"b" == "a\\"b"', '"x\\\\" + REF != "x"', '0x10 == 16', '!true || !false', '3 - 1 - 1 == 1', '2 * 3 + 4 == 10',
    '1 < 2 && 2.5 > 2 && 2 <= 2 && 3 >= 3.0', 'QUAL == "." && FILTER == "PASS"',
]
BAD_EXPRS = ['', 'REF ==', '(REF == "A"', 'REF = "A"', 'REF & "A"', '"abc', '"a\\n"', 'REF + 1 == "A1"', '!REF', 'REF && true',
             'POS < 5', 'len(REF) == 1', 'REF == "A", true', '9223372036854775807 + 1 == 0', '1 / 0 == 1']


def test_expression_subset_against_oracle_evaluator(pgb, tmp_path):
    prefix = str(tmp_path / "x")
    synth.make_pfile(prefix, 11, 40, 6, "lean")
    vh, vrows = onp.parse_table(open(prefix + ".pvar", "rb").read())
    for ex in EXPRS:
        want = onp.filter_metadata(vh, vrows, ex)
        got = pgb.VcfPlan(prefix, None, ex).var_idx.tolist()
        assert got == want, ex
    for ex in BAD_EXPRS:
        with pytest.raises(onp.OracleError):
            onp.filter_metadata(vh, vrows, ex)
        with pytest.raises(pgb.PgbError) as ei:
            pgb.VcfPlan(prefix, None, ex)
        assert ei.value.status == pgb.E_EXPR, ex


def test_query_subcommand_matches_oracle(pgb, basic1, tmp_path):
    # README example: pgen-rs query data/basic1/basic1 -i 'ALT == "G"' -f 'CHROM + " " + POS'
    r = subprocess.run([CLI, "query", basic1, "-i", 'ALT == "G"', "-f", 'CHROM + " " + POS'], capture_output=True)
    assert r.returncode == 0, r.stderr
    vh, vrows = onp.parse_table(open(basic1 + ".pvar", "rb").read())
    want = onp.query_metadata(vh, vrows, 'ALT == "G"', 'CHROM + " " + POS')
    assert r.stdout.decode().splitlines() == want and len(want) == 4130
    # -s queries the samples; long flags and --flag=value spellings (clap)
    r = subprocess.run([CLI, "query", "--samples", "--fstring=IID", "--include", 'IID >= "NA20900"', basic1], capture_output=True)
    sh, srows = onp.parse_table(open(basic1 + ".psam", "rb").read())
    assert r.stdout.decode().splitlines() == onp.query_metadata(sh, srows, 'IID >= "NA20900"', "IID")
    # through the C ABI
    out = tmp_path / "q.txt"
    fd = os.open(out, os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
    pgb.pfile_query(basic1, 'ID', 'REF == "T"', False, fd)
    os.close(fd)
    assert out.read_text().splitlines() == onp.query_metadata(vh, vrows, 'REF == "T"', "ID")


def test_cli_surface(tmp_path):
    r = subprocess.run([CLI, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "query" in r.stdout and "filter" in r.stdout
    r = subprocess.run([CLI, "filter", "--help"], capture_output=True, text=True)
    for flag in ("--include-var", "--include-sam", "-o, --out"):
        assert flag in r.stdout
    r = subprocess.run([CLI, "query", "--help"], capture_output=True, text=True)
    for flag in ("-f, --fstring", "-i, --include", "-s, --samples"):
        assert flag in r.stdout
    assert subprocess.run([CLI, "--version"], capture_output=True, text=True).stdout.startswith("pgen-b200 ")
    assert subprocess.run([CLI, "bogus"], capture_output=True).returncode == 2
    assert subprocess.run([CLI, "query", "p"], capture_output=True).returncode == 2  # --fstring is required
    # a missing pfile is the reference's File::open(...).unwrap() panic: exit 101
    assert subprocess.run([CLI, "filter", str(tmp_path / "nothing")], capture_output=True).returncode == 101
    assert subprocess.run([CLI, "query", str(tmp_path / "nothing"), "-f", "ID"], capture_output=True).returncode == 101


def _write_pgen10(path, m, n, lens, types, type_bits=4, len_bytes=1):
    B = (m + 65535) // 65536
    mode = (0 if type_bits == 4 else 4) + (len_bytes - 1)
    hdr = b"\x6c\x1b\x10" + m.to_bytes(4, "little") + n.to_bytes(4, "little") + bytes([mode | 0x40])
    body = b""
    for b in range(B):
        a, e = b * 65536, min(m, (b + 1) * 65536)
        t = types[a:e]
        if type_bits == 4:
            t = list(t) + [0] * (len(t) % 2)
            body += bytes(int(t[i]) | (int(t[i + 1]) << 4) for i in range(0, len(t), 2))
        else:
            body += bytes(int(x) for x in t)
        body += b"".join(int(x).to_bytes(len_bytes, "little") for x in lens[a:e])
    rec0 = 12 + 8 * B + len(body)
    offs, pos = [], rec0
    for b in range(B):
        offs.append(pos)
        pos += int(sum(lens[b * 65536:(b + 1) * 65536]))
    with open(path, "wb") as f:
        f.write(hdr + b"".join(o.to_bytes(8, "little") for o in offs) + body + bytes(pos - rec0))
    return rec0


@pytest.mark.parametrize("m,type_bits,len_bytes", [(5, 4, 1), (7, 8, 2), (65536, 4, 2), (65536 + 3, 4, 3), (131072, 8, 1)])
def test_pgen10_index(pgb, tmp_path, m, type_bits, len_bytes):
    """Mode-0x10 header geometry of src/pgen.rs:100-137 plus the per-variant record index,
    including the M % 65536 == 0 case the reference walker gets wrong (pgen.rs:200-204)."""
    rng = np.random.default_rng(m)
    lens = rng.integers(1, min(200, 2 ** (8 * len_bytes)), size=m)
    types = rng.integers(0, 16 if type_bits == 4 else 256, size=m)
    path = str(tmp_path / "s.pgen")
    rec0 = _write_pgen10(path, m, 10, lens, types, type_bits, len_bytes)
    info, off, typ, ln = pgb.pgen10_index(path)
    B = (m + 65535) // 65536
    assert (info.n_variants, info.n_samples, info.storage_mode) == (m, 10, 0x10)
    assert (info.record_type_bits, info.record_length_bytes, info.allele_count_bytes, info.provisional_ref_storage) == (type_bits, len_bytes, 0, 1)
    assert info.variant_block_count == B and info.main_header_body_offset == 12 + 8 * B
    assert info.variant_records_offset == rec0
    assert (typ == types).all() and (ln == lens).all()
    assert (off == rec0 + np.concatenate([[0], np.cumsum(lens)])).all()
    assert os.path.getsize(path) == int(off[-1])
    # a fixed-width file is not mode 0x10
    p2 = tmp_path / "f.pgen"
    p2.write_bytes(synth.pgen_header(1, 4) + b"\0")
    with pytest.raises(pgb.PgbError) as ei:
        pgb.pgen10_index(str(p2))
    assert ei.value.status == pgb.E_MODE


def test_parallel_planning_equals_serial(pgb, tmp_path, monkeypatch):
    """The .pvar/.psam parser, the predicate filter and the prefix builder split large tables over
    worker threads at line boundaries; forced onto small random tables (mixed \\n / \\r\\n terminators,
    empty lines, empty fields) the result must equal the serial one and the numpy oracle's."""
    import synth
    rng = np.random.default_rng(8)
    for trial in range(12):
        n_rows = int(rng.integers(1, 400))
        n_cols = int(rng.integers(1, 7))
        cols = ["ID"] + ["C%d" % i for i in range(1, n_cols)]
        term = [b"\n", b"\r\n"][trial % 2]
        lines = [b"##meta=1" + term, ("#" + "\t".join(cols)).encode() + term]
        for r in range(n_rows):
            fields = ["" if rng.integers(0, 9) == 0 else "".join(chr(rng.integers(65, 91)) for _ in range(rng.integers(1, 12)))
                      for _ in range(n_cols)]
            fields[0] = "v%d" % (r % 7)
            lines.append("\t".join(fields).encode() + term)
            if rng.integers(0, 10) == 0:
                lines.append(term)  # empty line: skipped by csv
        prefix = str(tmp_path / ("t%d" % trial))
        open(prefix + ".pvar", "wb").write(b"".join(lines))
        open(prefix + ".psam", "wb").write(b"#IID\tSEX\ns1\t1\ns2\t2\n")
        open(prefix + ".pgen", "wb").write(synth.pgen_header(n_rows, 2) + bytes(n_rows))
        q = 'ID == "v3" || ID == "v5"' if trial % 3 else None
        plans = []
        for par_min, threads in (("1000000000", "1"), ("0", "5"), ("0", "3")):
            monkeypatch.setenv("PGB_HOST_PAR_MIN_BYTES", par_min)
            monkeypatch.setenv("PGB_HOST_THREADS", threads)
            plans.append(pgb.VcfPlan(prefix, None, q))
        for p in plans[1:]:
            assert (p.var_idx == plans[0].var_idx).all() and p.header == plans[0].header
            assert (p.prefix_off == plans[0].prefix_off).all() and (p.prefix_blob == plans[0].prefix_blob).all()
        vh, vrows = onp.parse_table(b"".join(lines))
        assert list(plans[0].var_idx) == onp.filter_metadata(vh, vrows, q)
        want = b"".join(onp.line_prefix(vrows[i]) for i in plans[0].var_idx)
        assert plans[0].prefix_blob.tobytes() == want
