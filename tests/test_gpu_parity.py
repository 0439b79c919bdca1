"""Parity proper: the CUDA path, called through the C ABI, against the oracle — bit-exact.

Small/medium shapes are compared byte for byte with the oracle; the full BASELINE sizes are
checked with size-independent properties (decode -> re-encode round trip of every byte on the
device, closed-form sizes, sampled lines against the oracle)."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

import oracle_np as onp
import synth
from conftest import ROOT, orc_output_vcf, write_case

pytestmark = pytest.mark.gpu
CLI = os.path.join(ROOT, "bin", "pgen-b200")


def sha(b):
    return hashlib.sha256(b).hexdigest()


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    torch.cuda.set_device(0)
    return torch


def image_of(recs, n):
    return np.concatenate([np.frombuffer(synth.pgen_header(recs.shape[0], n), dtype=np.uint8), recs.reshape(-1)])


def random_prefixes(rng, count, lo, hi):
    pre = [bytes(rng.integers(33, 127, size=rng.integers(lo, hi + 1), dtype=np.uint8)) for _ in range(count)]
    blob = np.frombuffer(b"".join(pre) + b"\0", dtype=np.uint8).copy()
    off = np.zeros(count + 1, np.uint64)
    off[1:] = np.cumsum([len(x) for x in pre])
    return pre, blob, off


def test_device_present(pgb):
    assert pgb.lib.pgb_device_count() >= 1


def test_golden_vectors_through_pfile_api(pgb, kat_cases, tmp_path):
    for case in kat_cases:
        prefix = write_case(case, tmp_path)
        out = prefix + ".gpu.vcf"
        pgb.pfile_output_vcf(prefix, case["sam_query"], case["var_query"], out)
        assert open(out, "rb").read() == case["vcf"].encode(), case["name"]


# bits 16-30 steer the batch path (k2_batch.cuh): 0x1.... off, 0x2.... on for keep-all too, smaller batches / budgets,
# 0x1....... / 0x2....... one / two bulk-stored shared-memory images per warp, 0x3....... direct stores,
# 0x4....... the other number of input stages
@pytest.mark.parametrize("variant", [0x000, 0x010, 0x020, 0x112, 0x222, 0x1410, 0x2822, 0x10000, 0x20000, 0x10220000,
                                     0x02120000, 0x07020000, 0x20020000, 0x50420000, 0x30020000])
def test_random_shapes_bit_exact(pgb, variant, monkeypatch):
    monkeypatch.setenv("PGB_K2_VARIANT", str(variant))
    rng = np.random.default_rng(variant + 5)
    sizes = [1, 2, 3, 4, 5, 7, 8, 15, 16, 17, 31, 33, 63, 64, 65, 100, 127, 129, 300, 511, 1000, 2504, 5000, 20000]
    for trial in range(int(os.environ.get("PGB_SOAK_TRIALS", "40"))):  # PGB_SOAK_TRIALS=2000 for a soak run
        n = int(rng.choice(sizes))
        m = int(rng.integers(1, 40))
        recs = rng.integers(0, 256, size=(m, synth.record_size(n)), dtype=np.uint8)
        mode = rng.integers(0, 3)
        sam = None if mode == 0 else np.sort(rng.choice(n, size=int(rng.integers(0, n + 1)), replace=False)).astype(np.uint32)
        var = None if rng.integers(0, 2) else np.sort(rng.choice(m, size=int(rng.integers(1, m + 1)), replace=False)).astype(np.uint32)
        rows = np.arange(m) if var is None else var
        lo, hi = [(0, 0), (0, 5), (1, 40), (30, 200)][rng.integers(0, 4)]
        pre, blob, off = random_prefixes(rng, len(rows), lo, hi)
        with pgb.PgenFile(image=image_of(recs, n)) as f:
            got = pgb.export_to_bytes(f, var, sam, blob, off)
        want = onp.format_body(recs, rows, np.arange(n) if sam is None else sam, pre)
        assert got == want, (variant, trial, n, m)


def test_basic1_config1_and_full_export(pgb, orc, basic1, tmp_path):
    """BASELINE configs[0] (filter) and the full export of data/basic1 (real metadata)."""
    out = str(tmp_path / "cfg1.vcf")
    st = pgb.pfile_output_vcf(basic1, 'IID == "NA20900"', 'ALT == "G"', out)
    assert (st.n_lines, st.n_kept_samples) == (4130, 1)
    want = str(tmp_path / "cfg1.want.vcf")
    onp.output_vcf(basic1, 'IID == "NA20900"', 'ALT == "G"', want)
    assert open(out, "rb").read() == open(want, "rb").read()
    assert os.path.getsize(out) == 718619
    out = str(tmp_path / "full.vcf")
    st = pgb.pfile_output_vcf(basic1, None, None, out)
    assert st.genotypes == 17784 * 2504 and st.kernel_launches > 0
    want = str(tmp_path / "full.want.vcf")
    assert orc_output_vcf(orc, basic1, None, None, want, 2) == 0
    assert os.path.getsize(out) == 181130024
    assert sha(open(out, "rb").read()) == sha(open(want, "rb").read())


def test_random1_full_export(pgb, orc, random1, tmp_path):
    """BASELINE configs[1]: 300 samples x 200 000 variants, R = 75 (every load phase)."""
    out = str(tmp_path / "r1.vcf")
    st = pgb.pfile_output_vcf(random1, None, None, out)
    assert st.genotypes == 60_000_000
    want = str(tmp_path / "r1.want.vcf")
    assert orc_output_vcf(orc, random1, None, None, want, 2) == 0
    assert sha(open(out, "rb").read()) == sha(open(want, "rb").read())


def test_cli_filter_end_to_end(orc, basic1, tmp_path):
    out = str(tmp_path / "cli.vcf")
    r = subprocess.run([CLI, "filter", basic1, "--include-sam", 'IID == "NA20900"', "--include-var", 'ALT == "G"', "-o", out],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    want = str(tmp_path / "cli.want.vcf")
    onp.output_vcf(basic1, 'IID == "NA20900"', 'ALT == "G"', want)
    assert open(out, "rb").read() == open(want, "rb").read()
    # default output name, src/main.rs:121-122
    r = subprocess.run([CLI, "filter", basic1, "--include-var", 'ID == "rs8100066"'], capture_output=True, text=True)
    assert r.returncode == 0 and os.path.exists(basic1 + ".pgen-rs.vcf")


def test_chunking_slots_and_sinks(pgb, tmp_path, monkeypatch):
    """Many small chunks (slot reuse), and the three sink kinds: positional file, O_APPEND
    file and a pipe (ordered writes)."""
    monkeypatch.setenv("PGB_CHUNK_MB", "1")
    rng = np.random.default_rng(9)
    n, m = 1001, 6000
    recs = rng.integers(0, 256, size=(m, synth.record_size(n)), dtype=np.uint8)
    pre, blob, off = random_prefixes(rng, m, 20, 60)
    want = onp.format_body(recs, np.arange(m), np.arange(n), pre)
    p = str(tmp_path / "c.pgen")
    synth.write_pgen_bytes(p, recs, n)
    with pgb.PgenFile(p) as f:
        out = str(tmp_path / "pos.vcf")
        fd = os.open(out, os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
        os.write(fd, b"HEADER\n")
        st = f.export_gt_vcf(None, None, blob, off, fd)
        assert st.n_chunks > 10
        assert os.lseek(fd, 0, os.SEEK_CUR) == 7 + len(want)
        os.write(fd, b"TAIL")
        os.close(fd)
        assert open(out, "rb").read() == b"HEADER\n" + want + b"TAIL"
        # the same through the O_DIRECT output stage (4 KiB-aligned interiors straight from the page-locked slots,
        # ragged ends buffered); falls back silently where the file system refuses O_DIRECT
        monkeypatch.setenv("PGB_ODIRECT", "1")
        for hdr in (b"HEADER\n", b"", b"x" * 4096, b"y" * 5000):
            out = str(tmp_path / "direct.vcf")
            fd = os.open(out, os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
            os.write(fd, hdr)
            f.export_gt_vcf(None, None, blob, off, fd)
            assert os.lseek(fd, 0, os.SEEK_CUR) == len(hdr) + len(want)
            os.close(fd)
            assert open(out, "rb").read() == hdr + want
        monkeypatch.delenv("PGB_ODIRECT")
        out = str(tmp_path / "app.vcf")
        open(out, "wb").write(b"H\n")
        fd = os.open(out, os.O_WRONLY | os.O_APPEND)
        f.export_gt_vcf(None, None, blob, off, fd)
        os.close(fd)
        assert open(out, "rb").read() == b"H\n" + want
        r, w = os.pipe()
        import threading
        got = []
        t = threading.Thread(target=lambda: got.append(os.fdopen(r, "rb").read()))
        t.start()
        f.export_gt_vcf(None, None, blob, off, w)
        os.close(w)
        t.join()
        assert got[0] == want
        # sparse variant selection (per-row staging) and the gather path from a file
        var = np.sort(rng.choice(m, size=37, replace=False)).astype(np.uint32)
        sam = np.sort(rng.choice(n, size=100, replace=False)).astype(np.uint32)
        pre2, blob2, off2 = random_prefixes(rng, len(var), 5, 30)
        assert pgb.export_to_bytes(f, var, sam, blob2, off2) == onp.format_body(recs, var, sam, pre2)


def test_argument_errors(pgb):
    rng = np.random.default_rng(3)
    n, m = 9, 4
    recs = rng.integers(0, 256, size=(m, 3), dtype=np.uint8)
    pre, blob, off = random_prefixes(rng, 2, 3, 3)
    with pgb.PgenFile(image=image_of(recs, n)) as f:
        def status(var, sam):
            try:
                pgb.export_to_bytes(f, var, sam, blob, off)
                return 0
            except pgb.PgbError as e:
                return e.status
        assert status([0, 3], [0, 8]) == 0
        assert status([0, 4], None) == pgb.E_RANGE          # row past the file: read_exact, pfile.rs:170
        assert status([0, 1], [0, 12]) == pgb.E_RANGE       # byte past the record: pfile.rs:173
        assert status([0, 1], [0, 11]) == 0                 # padding bits are readable in the reference
        assert status([0, 1], [3, 2]) == pgb.E_ARG          # never produced by filter_metadata
        assert status([0, 1], [2, 2]) == pgb.E_ARG
        out = np.zeros(4, np.uint8)
        with pytest.raises(pgb.PgbError) as ei:
            f.export_gt_vcf_mem([0, 1], None, blob, off, out.ctypes.data, 4)
        assert ei.value.status == pgb.E_SPACE
        with pytest.raises(pgb.PgbError) as ei:
            big = np.zeros(4096, np.uint8)
            f.export_gt_vcf_mem([0, 1], None, blob, off, big.ctypes.data, big.nbytes, devices=[99])
        assert ei.value.status == pgb.E_NO_DEVICE


def test_k0_k1_synth_device_level(pgb, torch_cuda):
    torch = torch_cuda
    rng = np.random.default_rng(12)
    # K0: keep-mask -> ascending index list
    for n in [1, 31, 32, 33, 1023, 1024, 1025, 2504, 70001]:
        for dens in [0.0, 0.1, 0.5, 1.0]:
            mask = (rng.random(n) < dens).astype(np.uint8) * rng.integers(1, 255, size=n).astype(np.uint8)
            d_mask = torch.from_numpy(mask).cuda()
            d_idx = torch.full((n + 8,), 0xFFFF, dtype=torch.int32, device="cuda")
            d_cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
            assert pgb.lib.pgb_dev_compact_samples(d_mask.data_ptr(), n, d_idx.data_ptr(), d_cnt.data_ptr(), 0) == 0
            torch.cuda.synchronize()
            want = np.nonzero(mask)[0]
            k = int(d_cnt.item())
            assert k == len(want)
            got = d_idx.cpu().numpy()
            assert (got[:k] == want).all() and (got[k:k + 8] == 0).all()
    # K1: record index + exclusive prefix sum of line lengths
    for n_lines in [1, 2, 255, 2048, 2049, 100_000]:
        plen = rng.integers(0, 300, size=n_lines).astype(np.uint64)
        base = 1000
        off = np.concatenate([[base], base + np.cumsum(plen)]).astype(np.uint64)
        rows = np.sort(rng.choice(4 * n_lines, size=n_lines, replace=False)).astype(np.uint32)
        kk, pitch = 77, 640
        d_off = torch.from_numpy(off.view(np.int64)).cuda()
        d_rows = torch.from_numpy(rows.view(np.int32)).cuda()
        d_meta = torch.zeros((n_lines + 1) * 4, dtype=torch.int64, device="cuda")
        d_scr = torch.zeros(pgb.lib.pgb_dev_index_scratch_bytes(n_lines) // 8 + 1, dtype=torch.int64, device="cuda")
        assert pgb.lib.pgb_dev_index_lines(d_rows.data_ptr(), d_off.data_ptr(), base, n_lines, kk, pitch, d_meta.data_ptr(),
                                           d_scr.data_ptr(), 0) == 0
        torch.cuda.synchronize()
        meta = d_meta.cpu().numpy().view(np.uint64).reshape(n_lines + 1, 4)
        lens = plen + 4 * kk + 1
        assert (meta[:, 0] == np.concatenate([[0], np.cumsum(lens)])).all()
        assert (meta[:-1, 1] == rows.astype(np.uint64) * pitch).all()
        assert (meta[:-1, 2] == off[:-1] - base).all()
        assert ((meta[:-1, 3] & 0xFFFFFFFF) == plen).all()
    # synthetic generators: device == numpy
    for n, rows0, nrows, pitch_pad in [(2504, 0, 300, 0), (5, 7, 64, 3), (300, 100000, 257, 5), (70001, 3, 9, 0),
                                       (500000, 199990, 10, 0), (6, 0, 33, 2)]:
        r = synth.record_size(n)
        pitch = r + pitch_pad
        d = torch.zeros(nrows * pitch, dtype=torch.uint8, device="cuda")
        assert pgb.lib.pgb_dev_synth_records(d.data_ptr(), pitch, 42, rows0, nrows, n, 0) == 0
        torch.cuda.synchronize()
        got = d.cpu().numpy().reshape(nrows, pitch)[:, :r]
        assert (got == synth.synth_records(42, rows0, nrows, n)).all()
        d.zero_()
        assert pgb.lib.pgb_dev_synth_records_fast(d.data_ptr(), pitch, 42, rows0, nrows, n, 0) == 0
        torch.cuda.synchronize()
        full = d.cpu().numpy().reshape(nrows, pitch)
        assert (full[:, :r] == synth.synth_records_fast(42, rows0, nrows, n)).all() and (full[:, r:] == 0).all()


def _device_run(pgb, torch, recs_dev, pitch, n, var_rows, kidx_np, blob, off, max_pfx, variant=0):
    n_lines = len(off) - 1
    k = n if kidx_np is None else len(kidx_np)
    total = int(off[-1] - off[0]) + n_lines * (4 * k + 1)
    d_blob = torch.from_numpy(blob).cuda()
    d_off = torch.from_numpy(off.view(np.int64)).cuda()
    d_rows = None if var_rows is None else torch.from_numpy(np.ascontiguousarray(var_rows, dtype=np.uint32).view(np.int32)).cuda()
    d_kidx = None if kidx_np is None else torch.from_numpy(np.concatenate([kidx_np, np.zeros(8, np.uint32)]).astype(np.uint32).view(np.int32)).cuda()
    d_meta = torch.zeros((n_lines + 1) * 4, dtype=torch.int64, device="cuda")
    d_scr = torch.zeros(pgb.lib.pgb_dev_index_scratch_bytes(n_lines) // 8 + 1, dtype=torch.int64, device="cuda")
    d_out = torch.full((total + 1024,), 0xAA, dtype=torch.uint8, device="cuda")
    pgb.dev_format(recs_dev, pitch, d_rows, d_blob, d_off, d_kidx, k, max_pfx, d_out, d_meta, d_scr, variant,
                   torch.cuda.current_stream().cuda_stream, int(off[0]))
    torch.cuda.synchronize()
    return d_out, total


def test_chr22_shape_full_size_properties(pgb, torch_cuda):
    """BASELINE configs[2] at full size (2 504 x 1 100 000, keep all): records synthesised on
    the device, formatted device-resident, then every output byte is checked by a round trip
    computed with plain torch ops (text -> codes -> re-packed bytes == the input records), and
    500 sampled lines are compared with the oracle."""
    torch = torch_cuda
    n, m, width = 2504, 1_100_000, 40
    r = synth.record_size(n)
    recs = torch.zeros(m * r + 64, dtype=torch.uint8, device="cuda")
    assert pgb.lib.pgb_dev_synth_records(recs.data_ptr(), r, 3, 0, m, n, 0) == 0
    blob, off = synth.uniform_prefix_blob(m, 0, width)
    d_out, total = _device_run(pgb, torch, recs, r, n, None, None, blob, off, width)
    L = width + 4 * n + 1
    assert total == m * L == 11_062_700_000
    assert bool((d_out[total:] == 0xAA).all()), "wrote past the end of the body"
    lines = d_out[:total].view(m, L)
    d_blob = torch.from_numpy(blob).cuda().view(m, width)
    step = 50_000
    lut = torch.full((256, 256), 255, dtype=torch.uint8, device="cuda")
    for code, txt in enumerate([b"00", b"01", b"11", b".."]):
        lut[txt[0], txt[1]] = code
    shifts = torch.tensor([0, 2, 4, 6], dtype=torch.uint8, device="cuda")
    for a in range(0, m, step):
        blk = lines[a:a + step]
        assert torch.equal(blk[:, :width], d_blob[a:a + step])
        assert bool((blk[:, L - 1] == 10).all())
        gt = blk[:, width:width + 4 * n].reshape(-1, n, 4)
        assert bool((gt[:, :, 0] == 9).all()) and bool((gt[:, :, 2] == 47).all())
        codes = lut[gt[:, :, 1].long(), gt[:, :, 3].long()]
        assert bool((codes < 4).all())
        packed = (codes.view(-1, n // 4, 4) << shifts).sum(dim=2, dtype=torch.uint8)
        assert torch.equal(packed, recs[a * r:(a + blk.shape[0]) * r].view(-1, r))
    rng = np.random.default_rng(0)
    pick = np.sort(rng.choice(m, size=500, replace=False))
    host_recs = recs[:m * r].view(m, r)[torch.from_numpy(pick).cuda()].cpu().numpy()
    assert (host_recs == np.concatenate([synth.synth_records(3, int(v), 1, n) for v in pick])).all()
    pre = [bytes(blob[int(off[v]):int(off[v + 1])]) for v in pick]
    want = onp.format_body(host_recs, np.arange(len(pick)), np.arange(n), pre)
    got = lines[torch.from_numpy(pick).cuda()].cpu().numpy().tobytes()
    assert got == want


def test_chr22_gather_heavy_full_size(pgb, orc, torch_cuda, tmp_path):
    """BASELINE configs[3] at full size: K = 250 of 2 504 samples (seed 4a), 550 000 of
    1 100 000 variants (seed 4b); the whole 0.57 GB body is compared with the C oracle."""
    torch = torch_cuda
    n, m = 2504, 1_100_000
    r = synth.record_size(n)
    recs = torch.zeros(m * r + 64, dtype=torch.uint8, device="cuda")
    assert pgb.lib.pgb_dev_synth_records(recs.data_ptr(), r, 3, 0, m, n, 0) == 0
    torch.cuda.synchronize()
    pgen = str(tmp_path / "chr22.pgen")
    with open(pgen, "wb") as f:
        f.write(synth.pgen_header(m, n))
        f.write(recs[:m * r].cpu().numpy().tobytes())
    sam = synth.subset_indices(41, n, 250)
    var = synth.subset_indices(42, m, 550_000)
    blob, off = synth.uniform_prefix_blob(m, 0, 48)
    blob = blob.reshape(m, 48)[var].reshape(-1).copy()
    off = (np.arange(len(var) + 1, dtype=np.uint64) * np.uint64(48))
    want_path = str(tmp_path / "want.body")
    fd = os.open(want_path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
    rc = orc.orc_export_body(pgen.encode(), var.ctypes.data, len(var), sam.ctypes.data, len(sam), blob.ctypes.data,
                             off.ctypes.data, fd, 2)
    os.close(fd)
    assert rc == 0
    want = open(want_path, "rb").read()
    assert len(want) == 550_000 * (48 + 1000 + 1)
    with pgb.PgenFile(pgen) as f:
        got = pgb.export_to_bytes(f, var, sam, blob, off)
    assert sha(got) == sha(want)
    # the same through the device-resident entry points (K0 for the index list)
    mask = np.zeros(4 * r, np.uint8)
    mask[sam] = 1
    d_mask = torch.from_numpy(mask).cuda()
    d_idx = torch.zeros(4 * r + 8, dtype=torch.int32, device="cuda")
    d_cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    assert pgb.lib.pgb_dev_compact_samples(d_mask.data_ptr(), 4 * r, d_idx.data_ptr(), d_cnt.data_ptr(), 0) == 0
    kidx = d_idx.cpu().numpy().view(np.uint32)[:250]
    assert (kidx == sam).all()
    d_out, total = _device_run(pgb, torch, recs, r, n, var, kidx, blob, off, 48)
    assert sha(d_out[:total].cpu().numpy().tobytes()) == sha(want)


def test_biobank_shape_wide_lines(pgb, torch_cuda):
    """BASELINE configs[4] shape (500 000 samples, 2 MB lines spanning 123 tiles) on a block of
    variants, byte for byte against the oracle, keep-all and a 50 % gather."""
    torch = torch_cuda
    n, m = 500_000, 48
    r = synth.record_size(n)
    host = synth.synth_records(5, 1000, m, n)
    recs = torch.zeros(m * r + 64, dtype=torch.uint8, device="cuda")
    assert pgb.lib.pgb_dev_synth_records(recs.data_ptr(), r, 5, 1000, m, n, 0) == 0
    torch.cuda.synchronize()
    assert (recs[:m * r].cpu().numpy().reshape(m, r) == host).all()
    blob, off = synth.prefix_blob("1000g", range(1000, 1000 + m), 5)
    pre = [bytes(blob[int(off[i]):int(off[i + 1])]) for i in range(m)]
    maxp = max(len(p) for p in pre)
    for variant in (0, 0x2011):
        d_out, total = _device_run(pgb, torch, recs, r, n, None, None, blob, off, maxp, variant)
        want = onp.format_body(host, np.arange(m), np.arange(n), pre)
        assert len(want) == total and sha(d_out[:total].cpu().numpy().tobytes()) == sha(want)
        assert bool((d_out[total:] == 0xAA).all())
    sam = synth.subset_indices(51, n, 250_000)
    d_out, total = _device_run(pgb, torch, recs, r, n, None, sam, blob, off, maxp)
    assert sha(d_out[:total].cpu().numpy().tobytes()) == sha(onp.format_body(host, np.arange(m), sam, pre))
    # and through the host pipeline from a memory image
    with pgb.PgenFile(image=image_of(host, n)) as f:
        assert sha(pgb.export_to_bytes(f, None, None, blob, off)) == sha(want)


def test_offsets_beyond_4gib_use_u64(pgb, orc, tmp_path):
    """The reference's record offset wraps at 2^32 (pfile.rs:165, u32 multiply).  On a sparse
    > 4 GiB .pgen the GPU path must agree with the u64 oracle and differ from the faithful one."""
    n = 500_000
    r = synth.record_size(n)
    first, cnt = 34355, 10  # 34360 * 125000 >= 2^32
    m = first + cnt
    p = str(tmp_path / "big.pgen")
    recs = synth.synth_records(5, first, cnt, n)
    try:
        with open(p, "wb") as f:
            f.write(synth.pgen_header(m, n))
            f.truncate(12 + m * r)
            f.seek(12 + first * r)
            f.write(recs.tobytes())
            # the rows the wrapped offsets land on hold different data
            f.seek(12 + ((first + 5) * r) % 2**32)
            f.write(synth.synth_records(6, 0, 5, n).tobytes())
    except OSError:
        pytest.skip("no sparse-file support here")
    if os.stat(p).st_blocks * 512 > 64 * r:
        pytest.skip("file system did not keep the file sparse")
    var = np.arange(first, first + cnt, dtype=np.uint32)
    sam = synth.subset_indices(7, n, 1000)
    blob, off = synth.prefix_blob("lean", var, 5)
    outs = {}
    for name, flags in (("u64", 2), ("u32", 3)):
        q = str(tmp_path / (name + ".body"))
        fd = os.open(q, os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
        assert orc.orc_export_body(p.encode(), var.ctypes.data, cnt, sam.ctypes.data, len(sam), blob.ctypes.data,
                                   off.ctypes.data, fd, flags) == 0
        os.close(fd)
        outs[name] = open(q, "rb").read()
    assert outs["u64"] != outs["u32"]
    with pgb.PgenFile(p) as f:
        got = pgb.export_to_bytes(f, var, sam, blob, off)
    assert got == outs["u64"]


def test_multi_device_sharding_matches_single(pgb, monkeypatch):
    ndev = pgb.lib.pgb_device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs (variant-range sharding across devices)")
    monkeypatch.setenv("PGB_CHUNK_MB", "4")
    rng = np.random.default_rng(21)
    n, m = 2504, 9000
    recs = synth.synth_records(8, 0, m, n)
    pre, blob, off = random_prefixes(rng, m, 30, 180)
    want = onp.format_body(recs, np.arange(m), np.arange(n), pre)
    with pgb.PgenFile(image=image_of(recs, n)) as f:
        for devs in ([0, 1], list(range(ndev))):
            assert pgb.export_to_bytes(f, None, None, blob, off, devices=devs) == want


def test_concurrent_exports_share_device_buffers(pgb):
    """Two host threads exporting at once on one device: the process-wide device buffers are leased
    to one export at a time, so both must come out right; pgb_release_buffers() then frees them and
    the next export re-creates them."""
    import threading
    rng = np.random.default_rng(31)
    jobs = []
    for n, m in ((1001, 700), (2504, 300)):
        recs = rng.integers(0, 256, size=(m, synth.record_size(n)), dtype=np.uint8)
        pre, blob, off = random_prefixes(rng, m, 10, 80)
        sam = np.sort(rng.choice(n, size=n // 3, replace=False)).astype(np.uint32)
        jobs.append((recs, n, blob, off, sam, onp.format_body(recs, np.arange(m), sam, pre)))
    got = [None, None]

    def work(i):
        recs, n, blob, off, sam, _ = jobs[i]
        with pgb.PgenFile(image=image_of(recs, n)) as f:
            for _ in range(5):
                got[i] = pgb.export_to_bytes(f, None, sam, blob, off)

    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert got[0] == jobs[0][5] and got[1] == jobs[1][5]
    pgb.lib.pgb_release_buffers()
    work(0)
    assert got[0] == jobs[0][5]


def test_tiny_shapes_exhaustive_alignment(pgb):
    """Every sample count 1..21 crossed with kept-sample patterns and prefix lengths 0..17: the line
    start, the GT start and the line end fall on every byte phase, including lines too short to hold
    a single aligned 16-byte chunk."""
    rng = np.random.default_rng(77)
    for n in range(0, 22):  # n = 0: a file without samples (R = 0): every line is prefix + newline
        m = 19
        recs = rng.integers(0, 256, size=(m, synth.record_size(n)), dtype=np.uint8)
        pre = [bytes(rng.integers(33, 127, size=(i * 7 + n) % 18, dtype=np.uint8)) for i in range(m)]
        blob = np.frombuffer(b"".join(pre) + b"\0", dtype=np.uint8).copy()
        off = np.zeros(m + 1, np.uint64)
        off[1:] = np.cumsum([len(x) for x in pre])
        sels = [None, np.zeros(0, np.uint32)]
        if n:
            sels += [np.arange(0, n, 2, dtype=np.uint32), np.array([n - 1], np.uint32)]
        with pgb.PgenFile(image=image_of(recs, n)) as f:
            for sam in sels:
                got = pgb.export_to_bytes(f, None, sam, blob, off)
                want = onp.format_body(recs, np.arange(m), np.arange(n) if sam is None else sam, pre)
                assert got == want, (n, None if sam is None else sam.tolist())


def test_failed_export_leaves_the_device_buffers_usable(pgb, tmp_path):
    """An I/O failure in the sink (descriptor opened read-only) is PGB_E_IO; the cached per-device
    buffers must then serve the next export normally."""
    rng = np.random.default_rng(41)
    n, m = 777, 900
    recs = rng.integers(0, 256, size=(m, synth.record_size(n)), dtype=np.uint8)
    pre, blob, off = random_prefixes(rng, m, 5, 40)
    want = onp.format_body(recs, np.arange(m), np.arange(n), pre)
    ro = str(tmp_path / "ro.vcf")
    open(ro, "wb").close()
    with pgb.PgenFile(image=image_of(recs, n)) as f:
        for sink_flags in (os.O_RDONLY, os.O_RDONLY | os.O_APPEND):
            fd = os.open(ro, sink_flags)
            with pytest.raises(pgb.PgbError) as ei:
                f.export_gt_vcf(None, None, blob, off, fd)
            os.close(fd)
            assert ei.value.status == pgb.E_IO
            assert pgb.export_to_bytes(f, None, None, blob, off) == want


def test_variant_order_is_the_callers(pgb, monkeypatch):
    """var_idx is written in the order given (the reference always passes ascending rows, but the
    ABI does not require it): shuffled and descending selections, in one chunk and in many."""
    rng = np.random.default_rng(55)
    n, m = 1203, 3000
    recs = rng.integers(0, 256, size=(m, synth.record_size(n)), dtype=np.uint8)
    with pgb.PgenFile(image=image_of(recs, n)) as f:
        for chunk_mb in ("128", "1"):
            monkeypatch.setenv("PGB_CHUNK_MB", chunk_mb)
            for var in (rng.permutation(m).astype(np.uint32), np.arange(m - 1, -1, -3, dtype=np.uint32),
                        rng.choice(m, size=40, replace=False).astype(np.uint32)):
                pre, blob, off = random_prefixes(rng, len(var), 0, 50)
                sam = np.sort(rng.choice(n, size=200, replace=False)).astype(np.uint32)
                for sel in (None, sam):
                    got = pgb.export_to_bytes(f, var, sel, blob, off)
                    assert got == onp.format_body(recs, var, np.arange(n) if sel is None else sel, pre)


def test_prefix_longer_than_a_tile(pgb):
    """A .pvar row of tens of kilobytes (a huge INFO field): the prefix spans several tiles."""
    rng = np.random.default_rng(66)
    n, m = 301, 6
    recs = rng.integers(0, 256, size=(m, synth.record_size(n)), dtype=np.uint8)
    pre, blob, off = random_prefixes(rng, m, 9000, 40000)
    sam = np.sort(rng.choice(n, size=77, replace=False)).astype(np.uint32)
    with pgb.PgenFile(image=image_of(recs, n)) as f:
        for sel in (None, sam):
            got = pgb.export_to_bytes(f, None, sel, blob, off)
            assert got == onp.format_body(recs, np.arange(m), np.arange(n) if sel is None else sel, pre)


def _write_standard_pgen(path, n, types, lens, payload, type_bits=8, len_bytes=2):
    """A mode-0x10 .pgen: header-format byte, block-offset table, per-block record types and
    lengths (the layout src/pgen.rs:100-258 walks), then the records."""
    m = len(types)
    nb = (m + 65535) // 65536
    mode = (0 if type_bits == 4 else 4) + (len_bytes - 1)
    hdr = b"\x6c\x1b\x10" + m.to_bytes(4, "little") + n.to_bytes(4, "little") + bytes([mode | 0x40])
    body = b""
    for b in range(nb):
        a, e = b * 65536, min(m, (b + 1) * 65536)
        t = [int(x) for x in types[a:e]]
        if type_bits == 4:
            t = t + [0] * (len(t) % 2)
            body += bytes(t[i] | (t[i + 1] << 4) for i in range(0, len(t), 2))
        else:
            body += bytes(t)
        body += b"".join(int(x).to_bytes(len_bytes, "little") for x in lens[a:e])
    rec0 = 12 + 8 * nb + len(body)
    cum = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    offs = [rec0 + int(cum[b * 65536]) for b in range(nb)]
    with open(path, "wb") as f:
        f.write(hdr + b"".join(o.to_bytes(8, "little") for o in offs) + body + payload)


def test_standard_format_record_index(pgb, tmp_path):
    """EXTENSION: a mode-0x10 .pgen (two variant blocks) whose records are a mix of plain 2-bit
    records and compressed ones of random lengths.  The header walk feeds the device record index;
    kept plain records export bit-exactly (dense range DMA and compact staging), a kept compressed
    record is PGB_E_MODE, and pgb_open keeps rejecting the file like the reference (pfile.rs:53)."""
    rng = np.random.default_rng(100)
    n, m = 37, 65536 + 1500
    r = synth.record_size(n)
    plain = rng.random(m) < 0.8
    types = np.where(plain, 0, rng.integers(1, 8, size=m)).astype(np.uint8)
    lens = np.where(plain, r, rng.integers(1, 40, size=m)).astype(np.int64)
    payload = rng.integers(0, 256, size=int(lens.sum()) + 64, dtype=np.uint8).tobytes()
    path = str(tmp_path / "std.pgen")
    _write_standard_pgen(path, n, types, lens, payload)
    with pytest.raises(pgb.PgbError) as ei:
        pgb.PgenFile(path)
    assert ei.value.status == pgb.E_MODE
    plain_rows = np.nonzero(plain)[0].astype(np.uint32)
    sam = np.sort(rng.choice(n, size=11, replace=False)).astype(np.uint32)
    with pgb.PgenFile(path, standard=True) as f:
        assert (f.n_variants, f.n_samples, f.record_bytes) == (m, n, r)
        for rows in (plain_rows, plain_rows[::7], plain_rows[[3, 65000 % len(plain_rows), len(plain_rows) - 1]]):
            recs = onp.records_standard(path, rows)
            pre, blob, off = random_prefixes(rng, len(rows), 0, 30)
            for sel in (None, sam):
                got = pgb.export_to_bytes(f, rows, sel, blob, off)
                assert got == onp.format_body(recs, np.arange(len(rows)), np.arange(n) if sel is None else sel, pre)
        bad = np.array([plain_rows[0], np.nonzero(~plain)[0][0]], dtype=np.uint32)
        pre, blob, off = random_prefixes(rng, 2, 1, 5)
        with pytest.raises(pgb.PgbError) as ei:
            pgb.export_to_bytes(f, bad, None, blob, off)
        assert ei.value.status == pgb.E_MODE


def test_logical_shards_on_one_gpu_concatenate(pgb):
    """The multi-GPU partition exercised with one physical GPU ("fake cluster"): the
    contiguous ranges of pgb_shard_plan exported one after the other must concatenate to the single
    export, and each shard's size must equal the planned byte range."""
    rng = np.random.default_rng(91)
    n, m = 911, 5000
    recs = synth.synth_records(9, 0, m, n)
    var = np.sort(rng.choice(m, size=3500, replace=False)).astype(np.uint32)
    sam = np.sort(rng.choice(n, size=300, replace=False)).astype(np.uint32)
    pre, blob, off = random_prefixes(rng, len(var), 0, 120)
    with pgb.PgenFile(image=image_of(recs, n)) as f:
        whole = pgb.export_to_bytes(f, var, sam, blob, off)
        assert whole == onp.format_body(recs, var, sam, pre)
        for shards in (2, 4, 8):
            lb, bb = pgb.shard_plan(len(sam), off, shards)
            parts = []
            for g in range(shards):
                a, b = int(lb[g]), int(lb[g + 1])
                sub_off = off[a:b + 1]
                parts.append(pgb.export_to_bytes(f, var[a:b], sam, blob, sub_off) if b > a else b"")
                assert len(parts[-1]) == int(bb[g + 1] - bb[g])
            assert b"".join(parts) == whole


def test_prefixes_built_on_the_device_from_raw_pvar_rows(pgb, monkeypatch):
    """pgb_export_gt_vcf_rows: the line prefix is the row as it is in the .pvar image + "\\tGT" appended by K2
    (pfile.rs:157-161); dense selections DMA the covering text range, sparse ones pack the kept rows, and the
    caller's row order is kept.  Checked against the oracle's grammar on a made-up .pvar with \\n and \\r\\n
    terminators, empty lines and rows of 1 to 300 bytes."""
    rng = np.random.default_rng(123)
    n, m = 911, 4000
    recs = rng.integers(0, 256, size=(m, synth.record_size(n)), dtype=np.uint8)
    rows = [bytes(rng.integers(33, 127, size=int(rng.integers(1, 301)), dtype=np.uint8)) for _ in range(m)]
    text = b"##comment\n#CHROM\tPOS\n"
    off = np.zeros(m, np.uint64)
    for i, r in enumerate(rows):
        off[i] = len(text)
        text += r + [b"\n", b"\r\n", b"\n\n"][i % 3]
    ln = np.array([len(r) for r in rows], np.uint32)
    tx = np.frombuffer(text, dtype=np.uint8)
    sam = np.sort(rng.choice(n, size=150, replace=False)).astype(np.uint32)
    with pgb.PgenFile(image=image_of(recs, n)) as f:
        for chunk_mb in ("128", "1"):
            monkeypatch.setenv("PGB_CHUNK_MB", chunk_mb)
            for var in (np.arange(m, dtype=np.uint32), np.sort(rng.choice(m, size=m // 2, replace=False)).astype(np.uint32),
                        np.sort(rng.choice(m, size=9, replace=False)).astype(np.uint32), rng.permutation(m)[:700].astype(np.uint32),
                        np.zeros(0, np.uint32)):
                for sel in (None, sam, np.zeros(0, np.uint32)):
                    got = pgb.export_rows_to_bytes(f, var, sel, tx, off[var], ln[var])
                    want = onp.format_body(recs, var, np.arange(n) if sel is None else sel, [rows[v] + b"\tGT" for v in var])
                    assert got == want, (chunk_mb, len(var), None if sel is None else len(sel))
        # a row outside the image is refused
        with pytest.raises(pgb.PgbError) as ei:
            pgb.export_rows_to_bytes(f, [0], None, tx, np.array([len(text) - 2], np.uint64), np.array([5], np.uint32))
        assert ei.value.status == pgb.E_RANGE


def test_chunk_planner_binary_search_equals_linear_scan(pgb, monkeypatch):
    """Ascending selections (what filter_metadata produces) get their chunk boundaries by binary search; the chunks
    and the bytes must be the ones of the linear scan, for keep-all, subsets, both prefix modes and both chunk limits."""
    rng = np.random.default_rng(321)
    n, m = 777, 20000
    recs = rng.integers(0, 256, size=(m, synth.record_size(n)), dtype=np.uint8)
    rows = [bytes(rng.integers(33, 127, size=int(rng.integers(1, 120)), dtype=np.uint8)) for _ in range(m)]
    text = b"#h\n"
    off = np.zeros(m, np.uint64)
    for i, r in enumerate(rows):
        off[i] = len(text)
        text += r + b"\n"
    ln = np.array([len(r) for r in rows], np.uint32)
    tx = np.frombuffer(text, dtype=np.uint8)
    sam = np.sort(rng.choice(n, size=90, replace=False)).astype(np.uint32)
    most = 0
    with pgb.PgenFile(image=image_of(recs, n)) as f:
        for env in ({"PGB_CHUNK_MB": "2"}, {"PGB_CHUNK_MB": "64", "PGB_CHUNK_IN_MB": "1"}):
            for var in (np.arange(m, dtype=np.uint32), np.sort(rng.choice(m, size=m // 3, replace=False)).astype(np.uint32)):
                pre = [rows[v] + b"\tGT" for v in var]
                blob = np.frombuffer(b"".join(pre) + b"\0", dtype=np.uint8).copy()
                poff = np.zeros(len(var) + 1, np.uint64)
                poff[1:] = np.cumsum([len(x) for x in pre])
                for sel in (None, sam):
                    want = onp.format_body(recs, var, np.arange(n) if sel is None else sel, pre)
                    res = {}
                    for linear in ("0", "1"):
                        for k, v in env.items():
                            monkeypatch.setenv(k, v)
                        monkeypatch.setenv("PGB_PLAN_LINEAR", linear)
                        out = np.empty(len(want), np.uint8)
                        nb, st = f.export_gt_vcf_mem(var, sel, blob, poff, out.ctypes.data, len(want))
                        assert out.tobytes() == want
                        assert pgb.export_rows_to_bytes(f, var, sel, tx, off[var], ln[var]) == want
                        res[linear] = int(st.n_chunks)
                    assert res["0"] == res["1"], (env, len(var), res)
                    most = max(most, res["0"])
                for k in env:
                    monkeypatch.delenv(k)
    assert most > 8


@pytest.mark.parametrize("variant", [0, 0x10020000, 0x20020000, 0x40020000])
def test_batch_kernel_many_batches_per_cta(pgb, variant, monkeypatch):
    """The persistent batch kernel (k2_batch.cuh) cycling its shared-memory stages many times per CTA: tens of
    thousands of short lines with ragged prefixes; kept-sample counts on both sides of the register-plan limit
    (ceil(K/4) <= 64) and of one compaction round (32 bytes); keep-all short lines.  Every byte against the oracle."""
    monkeypatch.setenv("PGB_K2_VARIANT", str(variant))
    rng = np.random.default_rng(variant + 17)
    for n, k, m in ((2504, 97, 40000), (2504, 250, 30000), (2504, 300, 30000), (1000, 700, 12000), (300, None, 50000),
                    (64, None, 60000)):
        recs = rng.integers(0, 256, size=(m, synth.record_size(n)), dtype=np.uint8)
        lens = rng.integers(0, 81, size=m)
        pool = rng.integers(33, 127, size=int(lens.sum()) + 1, dtype=np.uint8)
        off = np.zeros(m + 1, np.uint64)
        off[1:] = np.cumsum(lens)
        pre = [pool[int(off[i]):int(off[i + 1])].tobytes() for i in range(m)]
        sam = None if k is None else np.sort(rng.choice(n, size=k, replace=False)).astype(np.uint32)
        with pgb.PgenFile(image=image_of(recs, n)) as f:
            got = pgb.export_to_bytes(f, None, sam, pool, off)
        want = onp.format_body(recs, np.arange(m), np.arange(n) if sam is None else sam, pre)
        assert sha(got) == sha(want), (variant, n, k, m)


def test_rows_mode_with_rows_longer_than_a_tile(pgb):
    """On-device prefix construction when a .pvar row (a huge INFO field) spans several K2 tiles: the appended
    "\\tGT" may straddle a tile boundary."""
    rng = np.random.default_rng(404)
    n, m = 301, 8
    recs = rng.integers(0, 256, size=(m, synth.record_size(n)), dtype=np.uint8)
    rows = [bytes(rng.integers(33, 127, size=int(sz), dtype=np.uint8))
            for sz in (16381, 16382, 16383, 16384, 9000, 40000, 32765, 5)]
    text = b"#h\n"
    off = np.zeros(m, np.uint64)
    for i, r in enumerate(rows):
        off[i] = len(text)
        text += r + b"\n"
    ln = np.array([len(r) for r in rows], np.uint32)
    tx = np.frombuffer(text, dtype=np.uint8)
    sam = np.sort(rng.choice(n, size=77, replace=False)).astype(np.uint32)
    with pgb.PgenFile(image=image_of(recs, n)) as f:
        for sel in (None, sam, np.zeros(0, np.uint32)):
            got = pgb.export_rows_to_bytes(f, None, sel, tx, off, ln)
            want = onp.format_body(recs, np.arange(m), np.arange(n) if sel is None else sel, [r + b"\tGT" for r in rows])
            assert got == want
