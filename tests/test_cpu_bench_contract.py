"""bench.py's reference arm runs on the host alone, so its side of the JSON contract can be checked here: one line, the keys the
driver reads, the CPU arm description, and silence from every rank but 0 under a multi-rank launch."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env_extra, *args):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    r = run({}, "--config", "c2", "--gpus", "1", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e", "extrapolated", "fcycles_per_solve"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "ms" and d["higher_is_better"] is False and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["metric"] == "ms per MG solve to rel. residual 1e-10" and "512x512x64" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["fcycles_per_solve"] == 4 and d["value"] > 0            # the seamount block converges in 4 F-cycles, as on the GPU


def test_reference_arm_is_silent_on_other_ranks():
    r = run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--config", "c5", "--gpus", "2")
    assert r.returncode == 0 and r.stdout.strip() == ""
