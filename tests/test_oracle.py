"""The oracle pinned: hand-derived golden vectors, two independent restatements (C and
numpy) against each other, the reference's shipped metadata with the sizes its survey
states, and the edge cases of SURVEY.md Appendix B."""
import hashlib
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import oracle_np as onp
import synth
from conftest import orc_output_vcf, write_case


def test_kat_c_oracle(orc, kat_cases, tmp_path):
    for case in kat_cases:
        prefix = write_case(case, tmp_path)
        out = prefix + ".vcf"
        for flags in (0, 2):  # faithful I/O and bulk I/O produce the same bytes
            assert orc_output_vcf(orc, prefix, case["var_idx"], case["sam_idx"], out, flags) == 0, case["name"]
            assert open(out, "rb").read() == case["vcf"].encode(), case["name"]


def test_kat_numpy_oracle_with_queries(kat_cases, tmp_path):
    for case in kat_cases:
        prefix = write_case(case, tmp_path)
        out = prefix + ".np.vcf"
        onp.output_vcf(prefix, case["sam_query"], case["var_query"], out)
        assert open(out, "rb").read() == case["vcf"].encode(), case["name"]
        vh, vrows = onp.parse_table(case["pvar"].encode())
        sh, srows = onp.parse_table(case["psam"].encode())
        assert onp.filter_metadata(vh, vrows, case["var_query"]) == case["var_idx"]
        assert onp.filter_metadata(sh, srows, case["sam_query"]) == case["sam_idx"]


def test_decode_table_by_hand(orc):
    # pfile.rs:172-183: byte 0b11_10_01_00 -> samples 0..3 = 0/0 0/1 1/1 ./.
    rec = np.array([0b11100100, 0b00011011], dtype=np.uint8)
    out = np.zeros(32, dtype=np.uint8)
    n = orc.orc_format_gt_fields(rec.ctypes.data, None, 0, 8, out.ctypes.data)
    assert out[:n].tobytes() == b"\t0/0\t0/1\t1/1\t./.\t./.\t1/1\t0/1\t0/0"
    assert onp.gt_text_block(onp.decode_matrix(rec[None, :], np.arange(8))).tobytes() == out[:n].tobytes()


def test_record_geometry(orc):
    # pfile.rs:196-200
    for n, r in [(0, 0), (1, 1), (3, 1), (4, 1), (5, 2), (300, 75), (2504, 626), (500000, 125000)]:
        assert orc.orc_record_size(n) == r == onp.record_size(n)
    # pfile.rs:165: u32 wrap from variant 34360 on at R = 125000 (cfg 5)
    assert orc.orc_record_offset(34359, 125000, 1) == 12 + 34359 * 125000
    assert orc.orc_record_offset(34360, 125000, 1) == 12 + (34360 * 125000) % 2**32
    assert orc.orc_record_offset(34360, 125000, 0) == 12 + 34360 * 125000
    assert onp.record_offset(34360, 125000, True) == orc.orc_record_offset(34360, 125000, 1)


def test_header_checks(orc, tmp_path):
    import ctypes
    m, n = ctypes.c_uint32(), ctypes.c_uint32()
    good = synth.pgen_header(3, 5) + bytes(6)
    cases = {b"": -1, good[:1]: -1, b"\x6c\x1c" + good[2:]: -2, good[:2] + b"\x10" + good[3:]: -3,
             good[:11] + b"\x00" + good[12:]: -4, good[:11]: -1, good: 0}
    for i, (blob, rc) in enumerate(cases.items()):
        p = tmp_path / f"h{i}.pgen"
        p.write_bytes(blob)
        assert orc.orc_read_pgen_header(str(p).encode(), ctypes.byref(m), ctypes.byref(n)) == rc
        if rc == 0:
            assert (m.value, n.value) == (3, 5) == onp.read_pgen_header(str(p))
        else:
            with pytest.raises(onp.OracleError):
                onp.read_pgen_header(str(p))
    assert orc.orc_read_pgen_header(str(tmp_path / "missing.pgen").encode(), ctypes.byref(m), ctypes.byref(n)) == -1


def test_basic1_config1_sizes_and_agreement(orc, basic1, tmp_path):
    """BASELINE config 1 on the reference's real metadata: Mk = 4130, K = 1 (sample 2444);
    the VCF is 11 866 header + 706 753 body bytes whatever the genotypes."""
    out_np = str(tmp_path / "np.vcf")
    onp.output_vcf(basic1, 'IID == "NA20900"', 'ALT == "G"', out_np)
    vh, vrows = onp.parse_table(open(basic1 + ".pvar", "rb").read())
    sh, srows = onp.parse_table(open(basic1 + ".psam", "rb").read())
    vi = onp.filter_metadata(vh, vrows, 'ALT == "G"')
    si = onp.filter_metadata(sh, srows, 'IID == "NA20900"')
    assert len(vrows) == 17784 and len(srows) == 2504  # data/basic1/basic1.log:15-17
    assert len(vi) == 4130 and si == [2444]
    out_c = str(tmp_path / "c.vcf")
    assert orc_output_vcf(orc, basic1, vi, si, out_c, 0) == 0
    a, b = open(out_np, "rb").read(), open(out_c, "rb").read()
    assert a == b
    assert len(a) == 718619
    header_end = a.index(b"\n#CHROM") + 1
    header_end = a.index(b"\n", header_end) + 1
    assert header_end == 11866


def test_basic1_full_export_size(orc, basic1, tmp_path):
    out_c = str(tmp_path / "full.vcf")
    assert orc_output_vcf(orc, basic1, None, None, out_c, 2) == 0
    assert os.path.getsize(out_c) == 181130024
    out_np = str(tmp_path / "full_np.vcf")
    onp.output_vcf(basic1, None, None, out_np)
    h = lambda p: hashlib.sha256(open(p, "rb").read()).hexdigest()
    assert h(out_c) == h(out_np)


@settings(max_examples=60, deadline=None)
@given(n=st.integers(1, 70), m=st.integers(1, 9), seed=st.integers(0, 2**31), data=st.data())
def test_c_vs_numpy_random(orc, tmp_path_factory, n, m, seed, data):
    d = tmp_path_factory.mktemp("rnd")
    prefix = str(d / "r")
    rng = np.random.default_rng(seed)
    recs = rng.integers(0, 256, size=(m, synth.record_size(n)), dtype=np.uint8)
    synth.write_pgen_bytes(prefix + ".pgen", recs, n)
    kind = data.draw(st.sampled_from(["random1", "lean", "1000g"]))
    ncomm = data.draw(st.integers(0, 3))
    synth.write_pvar(prefix + ".pvar", kind, m, seed, comments=[f"##c{i}=x" for i in range(ncomm)])
    synth.write_psam(prefix + ".psam", n, data.draw(st.sampled_from(["#IID\tSEX", "#FID\tIID", "#IID"])))
    vi = sorted(data.draw(st.sets(st.integers(0, m - 1))))
    si = sorted(data.draw(st.sets(st.integers(0, n - 1))))
    use_all_v, use_all_s = data.draw(st.booleans()), data.draw(st.booleans())
    out_c, out_np = prefix + ".c.vcf", prefix + ".np.vcf"
    assert orc_output_vcf(orc, prefix, None if use_all_v else vi, None if use_all_s else si, out_c,
                          data.draw(st.sampled_from([0, 2]))) == 0
    onp.output_vcf(prefix, None, None, out_np, var_idx=None if use_all_v else vi, sam_idx=None if use_all_s else si)
    assert open(out_c, "rb").read() == open(out_np, "rb").read()


def test_appendix_b_edges(orc, tmp_path):
    prefix = str(tmp_path / "e")
    recs = np.array([[0x1B, 0x02], [0x00, 0x01]], dtype=np.uint8)
    synth.write_pgen_bytes(prefix + ".pgen", recs, 5)
    # column line with trailing blanks + CR is trimmed; '##' lines keep their own endings (pfile.rs:144,211)
    open(prefix + ".pvar", "wb").write(b"##a=1\r\n#CHROM\tPOS\tID \r\n1\t2\tx\n1\t3\ty\n")
    open(prefix + ".psam", "wb").write(b"#IID\np0\np1\np2\np3\np4\n")
    out = prefix + ".vcf"
    assert orc_output_vcf(orc, prefix, None, [0, 4], out, 0) == 0
    got = open(out, "rb").read()
    assert got == (b"##fileformat=VCFv4.2\n##source=pgen-rs\n##a=1\r\n#CHROM\tPOS\tID\tFORMAT\tp0\tp4\n"
                   b"1\t2\tx\tGT\t./.\t1/1\n1\t3\ty\tGT\t0/0\t0/1\n")
    onp.output_vcf(prefix, None, None, prefix + ".np", sam_idx=[0, 4])
    assert open(prefix + ".np", "rb").read() == got
    # .psam without IID -> panic (pfile.rs:125)
    open(prefix + ".psam", "wb").write(b"#FID\tSEX\nf\t1\n")
    assert orc_output_vcf(orc, prefix, None, None, out, 0) == -6
    with pytest.raises(onp.OracleError):
        onp.output_vcf(prefix, None, None, prefix + ".np")
    # ragged row -> csv error; quote -> rejected
    open(prefix + ".psam", "wb").write(b"#IID\tSEX\np0\t1\np1\n")
    assert orc_output_vcf(orc, prefix, None, None, out, 0) == -7
    open(prefix + ".psam", "wb").write(b"#IID\tSEX\np0\t\"1\"\n")
    assert orc_output_vcf(orc, prefix, None, None, out, 0) == -8
    # variant beyond the file -> read_exact failure (pfile.rs:170); sample beyond the record -> index panic (:173)
    open(prefix + ".psam", "wb").write(b"#IID\n" + b"".join(b"p%d\n" % i for i in range(12)))
    open(prefix + ".pvar", "wb").write(b"#CHROM\tPOS\n1\t2\n1\t3\n1\t4\n")
    assert orc_output_vcf(orc, prefix, [2], [0], out, 0) == -9
    assert orc_output_vcf(orc, prefix, [0], [8], out, 0) == -9
    # a sample index inside the padding bits of the last byte is NOT an error in the reference
    assert orc_output_vcf(orc, prefix, [0], [7], out, 0) == 0
    assert open(out, "rb").read().endswith(b"1\t2\tGT\t0/0\n")
    # no '#' line at all in the .pvar -> header_lines.pop().unwrap() panics (pfile.rs:217)
    open(prefix + ".pvar", "wb").write(b"1\t2\n")
    assert orc_output_vcf(orc, prefix, None, None, out, 0) == -5


def test_synth_generator_is_deterministic_and_covers_all_codes():
    a = synth.synth_records(3, 100, 50, 2504)
    b = synth.synth_records(3, 0, 150, 2504)[100:]
    assert (a == b).all()
    codes = onp.decode_matrix(a, np.arange(2504))
    frac = np.bincount(codes.reshape(-1), minlength=4) / codes.size
    assert 0.005 < frac[3] < 0.015 and frac[0] > 0.4 and frac[1] > 0.1 and frac[2] > 0.02
    # padding bits are zero when N % 4 != 0
    c = synth.synth_records(3, 0, 20, 5)
    assert (c[:, 1] & 0xFC == 0).all()
    assert hashlib.sha256(synth.synth_records(1, 0, 4, 2504).tobytes()).hexdigest() == \
        hashlib.sha256(synth.synth_records(1, 0, 4, 2504).tobytes()).hexdigest()
