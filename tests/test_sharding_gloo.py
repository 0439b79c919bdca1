"""N > 1 on CPU: world_size-2 `gloo` test of the variant-range sharding.

The data path has no collective: each rank owns a contiguous range of kept variants
(pgb_shard_plan, the partition pgb_export_gt_vcf and bench.py use) and its VCF chunk lands at a
byte offset known in closed form.  Here two CPU processes each take their range from the plan,
stand in for the device with the oracle (the checker; no GPU in this container), and rank 0
verifies that the chunks concatenated in rank order are the single-process body and that the
offsets agree with the gathered chunk sizes.  The timing reduction bench.py uses (max over
ranks) is exercised on the same process group."""
import hashlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case():
    import synth
    rng = np.random.default_rng(77)
    n, m = 301, 900
    recs = synth.synth_records(9, 0, m, n)
    var = np.sort(rng.choice(m, size=700, replace=False)).astype(np.uint32)
    sam = np.sort(rng.choice(n, size=120, replace=False)).astype(np.uint32)
    pre = [bytes(rng.integers(33, 127, size=rng.integers(5, 200), dtype=np.uint8)) for _ in var]
    off = np.zeros(len(var) + 1, np.uint64)
    off[1:] = np.cumsum([len(p) for p in pre])
    return n, recs, var, sam, pre, off


def _worker(rank, world, port, q):
    for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools"), os.path.join(ROOT, "pgen-rs_b200", "python")):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist

    import oracle_np as onp
    import pgb200

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        n, recs, var, sam, pre, off = _case()
        lb, bb = pgb200.shard_plan(len(sam), off, world)
        a, b = int(lb[rank]), int(lb[rank + 1])
        chunk = onp.format_body(recs, var[a:b], sam, pre[a:b])
        # sizes of all shards -> exclusive scan must equal the plan's byte offsets (no collective on
        # the data path; this gather is the test's own check)
        sizes = [None] * world
        dist.all_gather_object(sizes, len(chunk))
        assert [int(x) for x in bb] == [0] + list(np.cumsum(sizes)), (bb, sizes)
        chunks = [None] * world
        dist.all_gather_object(chunks, chunk)
        # bench.py's timing rule: max over ranks
        t = torch.tensor([1.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert float(t.item()) == float(world)
        dist.barrier()
        if rank == 0:
            whole = onp.format_body(recs, var, sam, pre)
            q.put((hashlib.sha256(b"".join(chunks)).hexdigest() == hashlib.sha256(whole).hexdigest(),
                   [int(x) for x in lb]))
    finally:
        dist.destroy_process_group()


def test_two_rank_variant_range_sharding():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    ok, lb = q.get(timeout=10)
    assert ok
    assert lb[0] == 0 and lb[-1] == 700 and 0 < lb[1] < 700


@pytest.mark.parametrize("shards", [1, 2, 3, 8, 64])
def test_shard_plan_properties(pgb, shards):
    rng = np.random.default_rng(shards)
    for n_var in (0, 1, 5, 1000):
        plen = rng.integers(0, 300, size=n_var).astype(np.uint64)
        off = np.concatenate([[17], 17 + np.cumsum(plen)]).astype(np.uint64)
        k = int(rng.integers(0, 3000))
        lb, bb = pgb.shard_plan(k, off, shards)
        assert lb[0] == 0 and lb[-1] == n_var and (np.diff(lb.astype(np.int64)) >= 0).all()
        lens = plen + np.uint64(4 * k + 1)
        cum = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        assert (bb == cum[lb.astype(np.int64)]).all()
        if n_var >= 1000:  # balanced by bytes: no shard exceeds an equal share by more than one line
            share = int(cum[-1]) // shards
            assert (np.diff(bb.astype(np.int64)) <= share + int(lens.max())).all()
