"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`: the oracle with the reference's
per-cycle Galerkin products and LU, on host cores) prints ONE JSON line with the keys the driver reads, and the
argument parser accepts the driver's command lines."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    e = dict(os.environ, CUDA_VISIBLE_DEVICES="", **(env or {}))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         timeout=600, env=e, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


def test_reference_arm_prints_one_contract_line():
    d = run_bench("--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1", "--cpu-n", "64")
    assert d["impl"] == "reference" and d["metric"] == "vcycle_fine_grid_dof_per_s" and d["unit"] == "DOF/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] >= 1 and d["value"] > 0
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["scaling"] in ("strong", "weak") and "workload" in d["config"] and "sample" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert abs(d["ms_per_step"] * 1e-3 * d["value"] - 65 * 65) < 1e-6 * 65 * 65        # value = sample DOF / time


def test_reference_arm_on_a_non_zero_rank_exits_without_work():
    e = dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0", "--cpu-n", "32"], capture_output=True, text=True,
                         timeout=300, env=e, cwd=ROOT)
    assert out.returncode == 0 and not [l for l in out.stdout.splitlines() if l.startswith("{")]
