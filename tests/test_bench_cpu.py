"""bench.py contract on a box without a GPU: the reference arm (oracle port on the host cores) prints ONE JSON line with
the keys the driver reads; the product arm refuses to run (no CPU fallback); the report-only helpers are plain arithmetic."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, env=e)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run("--impl", "reference", "--steps", "2", "--warmup", "1", "--batch", "2", "--timesteps", "2", "--cpu-sample-iters", "1")
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "trajectories/s" and d["higher_is_better"] is True
    assert d["metric"] == "trajectories_per_sec_full_sampling_loop" and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "B=2/GPU" in d["config"]["workload"] and d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_without_work():
    r = _run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--batch", "2", "--timesteps", "2",
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box without a GPU")
def test_product_arm_fails_loudly_without_a_gpu():
    r = _run("--steps", "1", "--warmup", "0", "--no-modes", "--no-cpu-baseline")
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)


def test_batch1_roofline_arithmetic():
    import bench

    pk = dict(hbm=6453.1, tensor=1417.3, tensor_burst=1676.5, source="measured (MEASURED_PEAKS.json)")
    r = bench.b1_roofline(16.0, 100, 64_127_260, pk)
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] == 6453.1
    assert r["achieved"] == pytest.approx(64_127_260 * 100 / 16.0e-3 / 1e9)
    assert r["frac"] == pytest.approx(r["achieved"] / 6453.1)


def test_config0_cpu_figure_is_a_whole_plan_at_batch_one():
    import bench

    d = bench.cpu_config0(reps=1)
    assert d["batch"] == 1 and d["T"] == 100 and d["sched"] == "guidance_ddpm" and d["ms_per_plan_p50"] > 0
    assert d["traj_per_s"] == pytest.approx(1e3 / d["ms_per_plan_p50"])
