"""bench.py contract pieces that run without a GPU: the `--impl reference` arm (the reference algorithm on the host
cores, a bounded sample) must print one JSON line with the driver's keys; non-zero ranks print nothing."""

import json
import os
import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent


def _run(extra_env=None, *args):
    env = dict(os.environ, **(extra_env or {}))
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "2",
                           "--warmup", "1", "--cpu-rows", "64", *args], capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "activations/sec" and d["unit"] == "activations/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("c1") and d["config"]["cpu_rows_per_step"] == 64
    cb = d["cpu_baseline"]
    # the reference's own modules when its package is staged (build container / oracle/_ref), else the oracle port
    from oracle import ref_harness

    assert cb["kind"] == ("reference" if ref_harness.reference_src() is not None else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and "64 rows" in cb["sample"]
    assert cb["fixed_ms_per_step"] >= 0 and cb["ms_per_row"] >= 0
    assert d["e2e"] == {"value": d["value"], "unit": "activations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_on_other_ranks_is_silent():
    r = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29999"},
             "--gpus", "2")
    assert r.returncode == 0, r.stderr[-2000:]
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
