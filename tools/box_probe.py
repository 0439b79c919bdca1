#!/usr/bin/env python
"""box_probe.py — what the GPU box's host side can do (printed as one JSON object):
cores, memory, file systems, whether O_DIRECT opens succeed, buffered vs O_DIRECT write
rates with 1/4/8 writers, host memory copy/fill bandwidth.  Used to give the end-to-end legs
of bench.py measured ceilings (profiles/README.md)."""
import json
import mmap
import os
import subprocess
import sys
import threading
import time

import numpy as np


def fs_of(path):
    try:
        out = subprocess.run(["df", "-T", path], capture_output=True, text=True).stdout.strip().splitlines()[-1].split()
        return {"dev": out[0], "type": out[1], "avail_gb": int(out[4]) / 1e6}
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)}


def write_rate(path, total, n_threads, direct, piece=8 << 20):
    flags = os.O_WRONLY | os.O_CREAT | os.O_TRUNC
    if direct:
        flags |= os.O_DIRECT
    try:
        fd = os.open(path, flags, 0o644)
    except OSError as e:
        return {"error": str(e)}
    try:
        os.ftruncate(fd, total)
        buf = mmap.mmap(-1, piece)  # page-aligned
        buf.write(b"\t0/1" * (piece // 4))
        per = total // n_threads // piece * piece
        err = []

        def work(t):
            try:
                o = t * per
                end = o + per
                while o < end:
                    o += os.pwrite(fd, buf, o)
            except OSError as e:
                err.append(str(e))

        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(t,)) for t in range(n_threads)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        dt = time.perf_counter() - t0
        if err:
            return {"error": err[0]}
        return {"gb_per_s": per * n_threads / dt / 1e9}
    finally:
        os.close(fd)
        try:
            os.unlink(path)
        except OSError:
            pass


def mem_bw(n_threads, nbytes=1 << 30):
    src = [np.ones(nbytes // n_threads, np.uint8) for _ in range(n_threads)]
    dst = [np.empty_like(s) for s in src]
    for d in dst:
        d[:] = 0

    def work(i):
        np.copyto(dst[i], src[i])

    best = 0
    for _ in range(3):
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(i,)) for i in range(n_threads)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        best = max(best, nbytes / (time.perf_counter() - t0) / 1e9)
    return best


def main():
    out = {"cores": os.cpu_count()}
    with open("/proc/meminfo") as f:
        mi = dict(line.split(":") for line in f)
    out["mem_total_gb"] = int(mi["MemTotal"].split()[0]) / 1e6
    out["mem_avail_gb"] = int(mi["MemAvailable"].split()[0]) / 1e6
    dirs = [d for d in ("/tmp", "/dev/shm", "/root", os.environ.get("GRAFT_REPO_ROOT", ".")) if os.path.isdir(d)]
    out["fs"] = {d: fs_of(d) for d in dirs}
    total = int(float(sys.argv[1]) * (1 << 30)) if len(sys.argv) > 1 else 4 << 30
    out["write"] = {}
    for d in dirs[:3]:
        r = {}
        for direct in (False, True):
            for nt in (1, 4, 8):
                r[("direct" if direct else "buffered") + "_%d" % nt] = write_rate(os.path.join(d, "pgb_probe.bin"), total, nt, direct)
        out["write"][d] = r
    out["host_copy_gb_per_s"] = {str(nt): mem_bw(nt) for nt in (1, 4, 8, 16)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
