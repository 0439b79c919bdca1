#!/usr/bin/env python
"""pin_against_ref.py — pins the oracle (and, with a GPU, the CUDA path) against the REAL `pgen-rs filter`.

TEST INFRASTRUCTURE.  Needs oracle/_ref/pgen-rs (oracle/build_ref.sh; a Rust toolchain and the crates of
Cargo.lock).  For every case it runs

    oracle/_ref/pgen-rs filter <prefix> [--include-var E] [--include-sam E] -o ref.vcf

and compares ref.vcf byte for byte with
  * the expected VCF of the hand-derived known-answer vectors (tests/golden/kat.json),
  * oracle/oracle_np.py (numpy restatement) and oracle/pgen_oracle.c (C restatement),
  * libpgb200 (pgb_pfile_output_vcf), when a CUDA device is present.

Cases: the KATs; data/basic1 (the reference's own .pvar/.psam, synthetic .pgen, seed 1) with BASELINE.json
configs[0]'s queries and without queries; the data/random1 shape (configs[1]).  Files > 4 GiB are NOT pinned this
way: there the reference itself is wrong (u32 record offset, src/pfile.rs:165) and parity is defined against the
u64 oracle.

Exit 0 = every case identical ("parity pinned"); 1 = a difference; 3 = no reference binary ("parity unpinned").
"""
import ctypes
import gzip
import hashlib
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools"), os.path.join(ROOT, "pgen-rs_b200", "python")):
    sys.path.insert(0, p)
REF_BIN = os.environ.get("PGB_REF_BIN", os.path.join(ROOT, "oracle", "_ref", "pgen-rs"))


def sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 22), b""):
            h.update(blk)
    return h.hexdigest()


def run_ref(prefix, sam_q, var_q, out):
    cmd = [REF_BIN, "filter", prefix]
    if var_q is not None:
        cmd += ["--include-var", var_q]
    if sam_q is not None:
        cmd += ["--include-sam", sam_q]
    cmd += ["-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return r.returncode, r.stderr


def main():
    if not os.path.exists(REF_BIN):
        print("parity unpinned: %s not built (oracle/build_ref.sh needs cargo + the crates of Cargo.lock)" % REF_BIN)
        return 3
    import oracle_np as onp
    import synth
    orc_so = os.path.join(ROOT, "oracle", "_build", "liborc.so")
    if not os.path.exists(orc_so):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    orc = ctypes.CDLL(orc_so)
    orc.orc_output_vcf.restype = ctypes.c_int
    orc.orc_output_vcf.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                   ctypes.c_char_p, ctypes.c_int]
    gpu = None
    try:
        import pgb200
        if pgb200.lib.pgb_device_count() > 0:
            gpu = pgb200
    except Exception:
        gpu = None
    bad = 0
    with tempfile.TemporaryDirectory(prefix="pgb_pin_") as td:
        cases = []
        with open(os.path.join(ROOT, "tests", "golden", "kat.json")) as f:
            for c in json.load(f)["cases"]:
                prefix = os.path.join(td, c["name"])
                open(prefix + ".pgen", "wb").write(bytes.fromhex(c["pgen_hex"].replace(" ", "")))
                open(prefix + ".pvar", "wb").write(c["pvar"].encode())
                open(prefix + ".psam", "wb").write(c["psam"].encode())
                cases.append((c["name"], prefix, c["sam_query"], c["var_query"], c["vcf"].encode()))
        data = os.path.join(ROOT, "tests", "data")
        b1 = os.path.join(td, "basic1")
        open(b1 + ".pvar", "wb").write(gzip.open(os.path.join(data, "basic1.pvar.gz")).read())
        open(b1 + ".psam", "wb").write(gzip.open(os.path.join(data, "basic1.psam.txt.gz")).read())
        synth.write_pgen(b1 + ".pgen", 1, 17784, 2504)
        cases.append(("basic1-config1", b1, 'IID == "NA20900"', 'ALT == "G"', None))
        cases.append(("basic1-full", b1, None, None, None))
        r1 = os.path.join(td, "random1")
        open(r1 + ".psam", "wb").write(gzip.open(os.path.join(data, "random1.psam.txt.gz")).read())
        synth.write_pvar(r1 + ".pvar", "random1", 200000, 2)
        synth.write_pgen(r1 + ".pgen", 2, 200000, 300)
        cases.append(("random1-full", r1, None, None, None))
        for name, prefix, sq, vq, want in cases:
            ref_out = prefix + ".ref.vcf"
            rc, err = run_ref(prefix, sq, vq, ref_out)
            if rc != 0:
                # the KATs include inputs the reference panics on; those carry no expected VCF
                print("%-28s reference exit %d (%s)" % (name, rc, err.strip().splitlines()[-1] if err.strip() else ""))
                if want is not None:
                    bad += 1
                continue
            ref_sha = sha(ref_out)
            res = []
            if want is not None:
                res.append(("kat", hashlib.sha256(want).hexdigest()))
            np_out = prefix + ".np.vcf"
            onp.output_vcf(prefix, sq, vq, np_out)
            res.append(("oracle_np", sha(np_out)))
            if sq is None and vq is None:
                c_out = prefix + ".c.vcf"
                if orc.orc_output_vcf(prefix.encode(), None, -1, None, -1, c_out.encode(), 0) == 0:
                    res.append(("pgen_oracle.c", sha(c_out)))
            if gpu is not None:
                g_out = prefix + ".gpu.vcf"
                gpu.pfile_output_vcf(prefix, sq, vq, g_out)
                res.append(("libpgb200", sha(g_out)))
            diff = [k for k, v in res if v != ref_sha]
            bad += len(diff)
            print("%-28s %s" % (name, "identical to pgen-rs: " + ", ".join(k for k, _ in res) if not diff else "DIFFERS: " + ", ".join(diff)))
    print("parity pinned by the reference binary" if not bad else "%d difference(s)" % bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
