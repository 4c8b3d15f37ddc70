"""bench.py contract (CPU): `--impl reference` prints exactly ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "hex elements assembled/s" and d["unit"] == "elements/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 1e3


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_reference_arm_same_config_and_threads_under_torchrun_env():
    """The reference arm runs on the GPU arm's workload description (same `config`) and ignores the OMP_NUM_THREADS=1 that
    torchrun exports (round-1 finding: the N > 1 CPU arm silently ran single-threaded)."""
    sys.path.insert(0, ROOT)
    import bench

    env = dict(os.environ, OMP_NUM_THREADS="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    d = json.loads(p.stdout.strip().splitlines()[-1])
    assert d["config"] == bench.workload_config(bench.ne_for(1), 1)
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["cpu_baseline"]["value_1thread"] > 1e3 and "element layers" in d["cpu_baseline"]["sample"]


def test_cpu_sample_is_a_slab_of_the_reference_mesh():
    """bench.host_mesh_slab(ne, L) == the first L element layers of the oracle's meshgrid + inflate_sphere."""
    sys.path.insert(0, ROOT)
    import numpy as np

    import bench
    from oracle import c_oracle, fem_oracle as o

    ne, L = 6, 2
    NL, IEN, ID, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
    o.inflate_sphere(NL, 0, 1, 0, 1)
    nl, ien, idd = bench.host_mesh_slab(ne, L)
    nN = (ne + 1) ** 2 * (L + 1)
    assert np.array_equal(nl, NL[:, :nN]) and np.array_equal(ien, IEN[: ne * ne * L]) and np.array_equal(idd, ID[:nN])
    # the slab sample assembles exactly the slab's elements: its K equals the sum of those element matrices in the full K's numbering
    Ks = c_oracle.assemble_system(ne, nl, ien, 3, "Q1", 3, idd, 40, 0.4, nthreads=2, nEl=ne * ne * L)
    assert Ks.m == 3 * nN and Ks.nnz == 9 * (3 * (ne + 1) - 2) ** 2 * (3 * (L + 1) - 2)
    rate, dt = bench.cpu_slab_rate(ne, L, 1)
    assert rate > 0 and dt > 0
