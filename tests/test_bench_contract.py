"""bench.py's output contract on CPU: the reference arm (`--impl reference`, cv2 on the host cores -- the only arm that runs
without a GPU) prints exactly ONE JSON line on stdout, whatever libraries write there, carrying the keys the driver reads;
under a multi-rank launch only rank 0 prints."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                          capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)


def test_reference_arm_prints_one_json_line():
    pytest.importorskip("cv2")
    p = _run({"RANK": "0", "WORLD_SIZE": "1", "LOCAL_RANK": "0"})
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, p.stdout[:2000]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "orb_extract_match_frames_per_s_640x480" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_are_silent():
    p = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert p.returncode == 0, p.stderr[-2000:]
    assert p.stdout.strip() == ""


def test_orbx_arm_refuses_to_run_without_a_gpu():
    """The product arm has no CPU fallback: on a box without a CUDA device it must stop with a clear message and print no JSON
    line (a line from a silent fallback would be read as a measurement)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True, cwd=ROOT, timeout=600,
                       env=dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0"))
    assert p.returncode != 0
    assert "no CPU fallback" in (p.stderr + p.stdout)
    assert not any(ln.strip().startswith("{") for ln in p.stdout.splitlines())
