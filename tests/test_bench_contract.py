"""bench.py contract checks that need no GPU: the reference arm (the oracle port timed on host
cores) prints one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_reference_arm_prints_contract_line():
    env = dict(os.environ, PGB_BENCH_REF_BUDGET_S="3")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--workload", "gather"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "genotypes/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "gather"


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1", PGB_BENCH_REF_BUDGET_S="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_host_side_ceiling_probes_run_without_a_gpu(tmp_path):
    """The host-side ceilings bench.py measures (raw sink strategies, CPU fill of the output buffer) are plain
    host code: exercise them on a small buffer so a broken probe cannot hide until a GPU run."""
    import sys
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    buf = np.full(8 << 20, 49, np.uint8)
    rates = bench.file_sink_ceilings(str(tmp_path / "probe.bin"), memoryview(buf), 4)
    assert set(rates) == {"buffered_pwrite_1", "buffered_pwrite_4", "o_direct_pwrite_4", "mmap_copy_4"}
    assert all(isinstance(v, float) and v > 0 for k, v in rates.items() if "o_direct" not in k)
    assert not (tmp_path / "probe.bin").exists()
    assert bench.host_write_probe(buf, 2) > 0


def test_config5_partition_covers_the_matrix_once():
    """Strong scaling of configs[4]: rank r owns variants [r*M/N, (r+1)*M/N) — contiguous, disjoint, complete."""
    m = 200_000
    for world in (1, 2, 3, 4, 8):
        ranges = [(r * m // world, (r + 1) * m // world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == m
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
