"""CPU: the parts of bench.py's contract that need no GPU — the reference arm runs on the host cores
and prints one JSON line with the agreed keys; under torchrun only rank 0 does."""
import json
import os
import subprocess
import sys

from conftest import ROOT

ARGS = ["--impl", "reference", "--steps", "2", "--warmup", "1", "--settle", "3", "--reference-budget-s", "3"]


def _run(extra_env=None, gpus="1"):
    env = dict(os.environ)
    env.update(extra_env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", gpus] + ARGS, env=env,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    return [l for l in out.stdout.splitlines() if l.strip()]


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    lines = _run()
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "particle-steps/sec" and d["unit"] == "particle-steps/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and abs(d["ms_per_step"] * 1e-3 * d["value"] - d["config"]["sample_particles"]) < 1e-3 * d["config"]["sample_particles"]
    assert d["config"]["workload"] == "dam-break-1M" and d["vs_baseline"] is None and d["dtype"] == "f32"
    # the same config keys as the GPU arm's line (workload, particles, h, dt, lattice, settle_steps), full-workload values
    assert d["config"]["particles"] == 1003520 and d["config"]["lattice"] == [64, 80, 196] and d["config"]["settle_steps"] == 3
    assert abs(d["config"]["h"] - 0.075) < 1e-9 and abs(d["config"]["dt"] - 0.0015) < 1e-7
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "particle slice" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_does_not_load_the_product():
    """The reference arm imports nothing of the product: its scene comes from the oracle's generator."""
    code = ("import sys, runpy; sys.argv = ['bench.py', '--gpus', '1'] + %r; "
            "runpy.run_path(%r, run_name='__main__'); "
            "bad = [m for m in sys.modules if m.startswith('sph_b200') or 'sph-fluid-simulator_b200' in m]; "
            "maps = open('/proc/self/maps').read(); "
            "assert not bad and 'libsph_b200' not in maps, (bad, 'libsph_b200' in maps)") % (ARGS, os.path.join(ROOT, "bench.py"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]


def test_reference_arm_runs_on_rank_zero_only():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, gpus="2") == []
    lines = _run({"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"}, gpus="2")
    assert len(lines) == 1 and json.loads(lines[0])["config"]["workload"].startswith("weak-scaling-8M-per-gpu")
