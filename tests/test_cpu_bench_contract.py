"""bench.py contract, checked on the CPU through the reference arm (the one arm that runs without a GPU):
exactly one JSON line on stdout, with the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

import harness as H


@pytest.mark.parametrize("config", ["c1", "c3", "m1"])
def test_reference_arm_prints_one_json_line(config):
    if H.load_ref() is None:
        pytest.skip("oracle/_ref not built")
    out = subprocess.run([sys.executable, os.path.join(H.ROOT, "bench.py"), "--impl", "reference", "--config", config,
                          "--npart", "60000", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"):
        assert key in d, key
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["value"] > 0 and d["unit"] == "pair_evals/s"


def test_non_zero_ranks_of_the_reference_arm_do_no_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(H.ROOT, "bench.py"), "--impl", "reference", "--config", "c1",
                          "--npart", "60000", "--gpus", "2"], capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
