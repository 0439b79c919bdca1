"""K2's per-lane byte/addressing logic (pgen-rs_b200/csrc/k2_core.cuh) run lane by lane on
the host and compared with the oracle: every output alignment phase, ragged sample counts,
gathers, empty selections, multi-tile lines.  This is a CPU pre-check of the kernel source;
the GPU parity tests are in test_gpu_parity.py."""
import numpy as np
import pytest

import oracle_np as onp
import synth

VARIANTS = [0x000, 0x010, 0x020, 0x001, 0x011, 0x021, 0x1000, 0x1010, 0x2020]


def run_sim(k2sim, rng, n, m, k_sel, var_sel, plen, phase, variant):
    r = synth.record_size(n)
    recs = rng.integers(0, 256, size=(m, r), dtype=np.uint8)
    flat = np.concatenate([recs.reshape(-1), np.zeros(32, np.uint8)])
    if var_sel is None:
        vr, vptr = np.arange(m, dtype=np.uint32), None
    else:
        vr = np.sort(rng.choice(m, size=var_sel, replace=False)).astype(np.uint32)
        vptr = vr.ctypes.data
    if k_sel is None:
        ki, sidx, k = None, np.arange(n), n
    else:
        sidx = np.sort(rng.choice(n, size=k_sel, replace=False)).astype(np.uint32)
        k = k_sel
        ki = np.concatenate([sidx, np.zeros(8, np.uint32)]).astype(np.uint32)
    pre = [bytes(rng.integers(33, 127, size=rng.integers(plen[0], plen[1] + 1), dtype=np.uint8)) for _ in vr]
    blob = np.frombuffer(b"".join(pre) + b"\0" * 8, dtype=np.uint8).copy()
    off = np.zeros(len(vr) + 1, np.uint64)
    off[1:] = np.cumsum([len(x) for x in pre])
    exp = onp.format_body(recs, vr, sidx, pre)
    guard = 1024
    buf = np.full(len(exp) + 2 * guard + 1024, 0xAA, np.uint8)
    start = guard + (-(buf.ctypes.data + guard)) % 512 + phase
    k2sim.sim_format_lines(flat.ctypes.data, r, vptr, len(vr), blob.ctypes.data, off.ctypes.data,
                           None if ki is None else ki.ctypes.data, k, buf.ctypes.data + start, variant)
    assert buf[start:start + len(exp)].tobytes() == exp
    assert (buf[:start] == 0xAA).all() and (buf[start + len(exp):] == 0xAA).all(), "wrote outside the body"


def test_every_phase_keep_all(k2sim):
    rng = np.random.default_rng(1)
    for phase in range(0, 64):
        run_sim(k2sim, rng, 301, 3, None, None, (0, 37), phase, VARIANTS[phase % len(VARIANTS)])


def test_every_phase_wide_prefix(k2sim):
    """Prefixes >= 64 bytes take the vectorised copy: every destination phase x source misalignment."""
    rng = np.random.default_rng(5)
    for phase in range(0, 64):
        run_sim(k2sim, rng, 77, 5, None if phase % 2 else 30, None, (64, 210), phase, VARIANTS[phase % len(VARIANTS)])


def test_every_phase_gather(k2sim):
    rng = np.random.default_rng(2)
    for phase in range(0, 64):
        run_sim(k2sim, rng, 301, 3, 97, 2, (5, 50), phase, VARIANTS[phase % len(VARIANTS)])


def test_random_shapes(k2sim):
    rng = np.random.default_rng(3)
    sizes = [1, 2, 3, 4, 5, 7, 8, 15, 16, 17, 31, 33, 63, 64, 65, 100, 127, 129, 300, 511, 1000, 2504, 5000]
    for _ in range(300):
        n = int(rng.choice(sizes))
        m = int(rng.integers(1, 10))
        k_sel = None if rng.integers(0, 3) == 0 else int(rng.integers(0, n + 1))
        var_sel = None if rng.integers(0, 2) else int(rng.integers(1, m + 1))
        plen = [(0, 0), (0, 5), (1, 40), (30, 200), (60, 70), (150, 400)][rng.integers(0, 6)]
        run_sim(k2sim, rng, n, m, k_sel, var_sel, plen, int(rng.integers(0, 512)), int(rng.choice(VARIANTS)))


@pytest.mark.parametrize("n", [20000, 70001])
def test_wide_lines_span_tiles(k2sim, n):
    rng = np.random.default_rng(4)
    for variant in (0, 0x1010, 0x2000, 0x1000):
        run_sim(k2sim, rng, n, 3, None, None, (10, 60), int(rng.integers(0, 512)), variant)
        run_sim(k2sim, rng, n, 3, n // 2, 2, (10, 60), int(rng.integers(0, 512)), variant)


def test_prefix_longer_than_a_tile(k2sim):
    """Prefixes of tens of kilobytes: the prefix itself spans several 4-16 KiB tiles and the GT text
    starts in a later tile than the line."""
    rng = np.random.default_rng(6)
    for variant in (0x1000, 0x0, 0x1010):
        run_sim(k2sim, rng, 301, 3, None, None, (9000, 40000), int(rng.integers(0, 512)), variant)
        run_sim(k2sim, rng, 2504, 2, 500, None, (17000, 20000), int(rng.integers(0, 512)), variant)


# ---- batch path (k2_batch.cuh) and the constant prefix suffix ("\tGT" appended on the device) ----

def run_sim_batch(k2sim, rng, n, m, k_sel, var_sel, plen, phase, B, sfx=b"", kidx_vec=1, per_line=False, variant=0,
                  unpacked=0):
    """per_line=True runs the per-line kernels (k2_core.cuh) with the suffix instead of the batch path."""
    r = synth.record_size(n)
    recs = rng.integers(0, 256, size=(m, r), dtype=np.uint8)
    # K2 reads whole 16-byte blocks around a record: keep the records 16-byte aligned inside a padded buffer
    store = np.zeros(m * r + 96, np.uint8)
    a0 = (-store.ctypes.data) % 16 + 16
    flat = store[a0:a0 + m * r + 32]
    flat[:m * r] = recs.reshape(-1)
    if var_sel is None:
        vr, vptr = np.arange(m, dtype=np.uint32), None
    else:
        vr = np.sort(rng.choice(m, size=var_sel, replace=False)).astype(np.uint32)
        vptr = vr.ctypes.data
    if k_sel is None:
        ki, sidx, k = None, np.arange(n), n
    else:
        sidx = np.sort(rng.choice(n, size=k_sel, replace=False)).astype(np.uint32)
        k = k_sel
        kstore = np.zeros(k + 8 + 8, np.uint32)
        s0 = ((-kstore.ctypes.data) % 16) // 4 + (0 if kidx_vec else 1)
        ki = kstore[s0:s0 + k + 8]
        ki[:k] = sidx
    pre = [bytes(rng.integers(33, 127, size=rng.integers(plen[0], plen[1] + 1), dtype=np.uint8)) for _ in vr]
    # the batch path fetches the 16-byte blocks around each prefix: keep 16 readable bytes on either side
    joined = b"".join(pre)
    bstore = np.zeros(len(joined) + 64, np.uint8)
    blob = bstore[16 + int(rng.integers(0, 16)):][:len(joined) + 8]
    blob[:len(joined)] = np.frombuffer(joined, dtype=np.uint8)
    off = np.zeros(len(vr) + 1, np.uint64)
    off[1:] = np.cumsum([len(x) for x in pre])
    exp = onp.format_body(recs, vr, sidx, [x + sfx for x in pre])
    guard = 1024
    buf = np.full(len(exp) + 2 * guard + 1024, 0xAA, np.uint8)
    start = guard + (-(buf.ctypes.data + guard)) % 512 + phase
    sfx_word = int.from_bytes(sfx.ljust(4, b"\0"), "little")
    if per_line:
        rc = k2sim.sim_format_lines_sfx(flat.ctypes.data, r, vptr, len(vr), blob.ctypes.data, off.ctypes.data,
                                        None if ki is None else ki.ctypes.data, k, buf.ctypes.data + start, variant,
                                        sfx_word, len(sfx))
    else:
        rc = k2sim.sim_format_lines_batch(flat.ctypes.data, r, r, vptr, len(vr), blob.ctypes.data, off.ctypes.data,
                                          None if ki is None else ki.ctypes.data, k, buf.ctypes.data + start, B,
                                          sfx_word, len(sfx), kidx_vec | (2 if unpacked else 0) | (4 if phase % 3 == 0 else 0) | (8 if phase % 5 < 2 else 0) | (16 if phase % 2 else 0))
    assert rc == 0
    got = buf[start:start + len(exp)].tobytes()
    if got != exp:
        bad = next(i for i in range(len(exp)) if got[i] != exp[i])
        raise AssertionError(f"first difference at byte {bad} of {len(exp)} (n={n} m={len(vr)} k={k} B={B} phase={phase})")
    assert (buf[:start] == 0xAA).all() and (buf[start + len(exp):] == 0xAA).all(), "wrote outside the body"


def test_batch_every_phase_gather(k2sim):
    rng = np.random.default_rng(11)
    for phase in range(0, 64):
        run_sim_batch(k2sim, rng, 301, 21, 97, 17, (5, 50), phase, [4, 5, 8, 32][phase % 4], kidx_vec=phase % 2)


def test_batch_every_phase_keep_all(k2sim):
    rng = np.random.default_rng(12)
    for phase in range(0, 64):
        run_sim_batch(k2sim, rng, 301, 13, None, None, (0, 37), phase, [4, 7, 16][phase % 3])


def test_batch_random_shapes(k2sim):
    rng = np.random.default_rng(13)
    sizes = [1, 2, 3, 4, 5, 7, 8, 15, 16, 17, 31, 33, 63, 64, 65, 100, 127, 129, 300, 511, 1000, 1025, 2504, 5000]
    for _ in range(400):
        n = int(rng.choice(sizes))
        m = int(rng.integers(1, 40))
        k_sel = None if rng.integers(0, 3) == 0 else int(rng.integers(0, n + 1))
        var_sel = None if rng.integers(0, 2) else int(rng.integers(1, m + 1))
        plen = [(0, 0), (0, 5), (1, 40), (30, 200), (60, 70), (150, 400)][rng.integers(0, 6)]
        sfx = [b"", b"\tGT", b"x", b"abcd"][rng.integers(0, 4)]
        run_sim_batch(k2sim, rng, n, m, k_sel, var_sel, plen, int(rng.integers(0, 512)), int(rng.integers(1, 33)), sfx,
                      kidx_vec=int(rng.integers(0, 2)), unpacked=int(rng.integers(0, 2)))


def test_batch_clustered_selection_stages_only_the_span(k2sim):
    """Kept samples confined to a slice of a wide record: only the covering byte span is staged."""
    rng = np.random.default_rng(14)
    n = 40000
    r = synth.record_size(n)
    recs = rng.integers(0, 256, size=(9, r), dtype=np.uint8)
    for lo, hi in ((17001, 17300), (0, 5), (39990, 40000), (20000, 20001)):
        sidx = np.sort(rng.choice(np.arange(lo, hi), size=min(hi - lo, 60), replace=False)).astype(np.uint32)
        ki = np.concatenate([sidx, np.zeros(8, np.uint32)]).astype(np.uint32)
        pre = [bytes(rng.integers(33, 127, size=rng.integers(3, 30), dtype=np.uint8)) for _ in range(9)]
        bstore = np.zeros(sum(len(x) for x in pre) + 64, np.uint8)
        blob = bstore[16:]
        blob[:len(b"".join(pre))] = np.frombuffer(b"".join(pre), dtype=np.uint8)
        off = np.zeros(10, np.uint64)
        off[1:] = np.cumsum([len(x) for x in pre])
        exp = onp.format_body(recs, np.arange(9), sidx, pre)
        store = np.zeros(9 * r + 96, np.uint8)
        a0 = (-store.ctypes.data) % 16 + 16
        flat = store[a0:a0 + 9 * r + 32]
        flat[:9 * r] = recs.reshape(-1)
        buf = np.full(len(exp) + 4096, 0xAA, np.uint8)
        start = 1024 + 3
        # R passed for the shared-memory sizing is the span bound the launcher would use (here: the full record)
        assert k2sim.sim_format_lines_batch(flat.ctypes.data, r, r, None, 9, blob.ctypes.data, off.ctypes.data, ki.ctypes.data,
                                            len(sidx), buf.ctypes.data + start, 4, 0, 0, 0) == 0
        assert buf[start:start + len(exp)].tobytes() == exp


def test_per_line_paths_with_suffix(k2sim):
    """The per-line kernels writing prefix = raw row bytes + constant suffix (on-device prefix construction)."""
    rng = np.random.default_rng(15)
    for trial in range(200):
        n = int(rng.choice([1, 5, 64, 301, 2504, 20000]))
        plen = [(0, 0), (0, 5), (1, 40), (60, 70), (150, 400)][rng.integers(0, 5)]
        k_sel = None if rng.integers(0, 2) else int(rng.integers(0, n + 1))
        sfx = [b"\tGT", b"x", b"abcd"][rng.integers(0, 3)]
        run_sim_batch(k2sim, rng, n, int(rng.integers(1, 6)), k_sel, None, plen, int(rng.integers(0, 512)), 0, sfx,
                      per_line=True, variant=int(rng.choice(VARIANTS)))
