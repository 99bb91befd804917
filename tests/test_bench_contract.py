"""The bench.py contract the round driver relies on, as far as it can be checked without a GPU:
the reference arm (`--impl reference`) times the CPU oracle and prints ONE JSON line with the
agreed keys; under a multi-rank launch only rank 0 works and prints."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(env_extra, *args):
    env = dict(os.environ)
    env.update(env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=env,
                          capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_one_contract_line():
    r = _run({}, "--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "cwt_scale_points_per_sec"
    assert d["unit"] == "scale-points/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] == 1 and d["warmup"] == 0
    assert d["n_gpus"] == 1 and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("config2") and d["config"]["n"] == 2 ** 20
    e2e = d["e2e"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert e2e["value"] == d["value"] and e2e["unit"] == d["unit"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]


def test_reference_arm_other_ranks_exit_quietly():
    r = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, "--impl", "reference", "--gpus", "2",
             "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_product_arm_fails_loudly_without_a_gpu():
    """No CPU fallback: without a device the product arm must not print a number."""
    import pycwt_b200._engine as eng
    try:
        have = eng.device_count() > 0
    except Exception:
        have = False
    if have:
        return
    r = _run({}, "--steps", "1", "--warmup", "0")
    assert r.returncode != 0
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
