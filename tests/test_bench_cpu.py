"""bench.py's CPU-runnable leg: the reference arm (`--impl reference`, the oracle port timed on the host cores) prints ONE
JSON line with the contract's keys.  The GPU arm needs a B200; its host-side bookkeeping (workload description, algorithmic
work per unit) is checked here without a device."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("config", ["wam", "mobile"])
def test_reference_arm_prints_one_contract_line(config):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", config, "--steps", "2",
                        "--warmup", "1", "--cpu-sample", "16", "--sdf", "60"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1                                  # ONE JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "trajectories/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and abs(d["value"] - 16 * 2 / (d["ms_per_step"] * 2e-3)) < 1e-6 * d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and "16 problems" in cb["sample"]
    assert "NOT upstream gpmp2/GTSAM" in cb["sample"]       # the baseline says what it is
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None and "workload" in d["config"] and "model" not in d["config"]
    assert d["reference_probe"]["found"] is False           # no gpmp2 + GTSAM build in this image


def test_reference_arm_other_ranks_exit_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_algorithmic_work_matches_the_survey_figures():
    """0.32 / 0.08 / 0.08 MFLOP per WAM linearization / solve / error evaluation, 976 lookups of 64 B."""
    sys.path.insert(0, ROOT)
    from gpmp2_b200 import synth
    cfg = synth.baseline_config("wam", sdf_cells=20)
    w = synth.algorithmic_work(cfg)
    assert abs(w["mflop_linearize"] - 0.32) < 0.01 and abs(w["mflop_solve"] - 0.08) < 0.005 and abs(w["mflop_error_eval"] - 0.08) < 0.005
    assert w["lookups_per_pass"] == 16 * 61 and w["l2_bytes_per_lookup"] == 64.0
    m = synth.algorithmic_work(synth.baseline_config("mobile"))
    assert m["lookups_per_pass"] == 10 * 61 and m["l2_bytes_per_lookup"] == 32.0      # 2-D quad cell: one 32-byte read
