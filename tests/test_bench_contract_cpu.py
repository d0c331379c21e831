"""The `bench.py --impl reference` arm (the CPU restatement timed on the host cores) on the tiny architecture: the JSON line
carries every key the driver's contract names, the same `config` object the GPU arm prints, and the e2e object of a run
without device copies.  (The GPU arm needs a B200; its line is checked by the driver and kept under profiles/.)"""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_the_contract_line(tmp_path):
    env = dict(**__import__("os").environ, DSOCR_BENCH_DIR=str(tmp_path))
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--config", "tiny", "--steps", "1",
                        "--warmup", "0", "--max-new-tokens", "8", "--cpu-tokens", "4"], capture_output=True, text=True,
                       env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]          # ONE json line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["metric"] == "pages/sec/box" and d["unit"] == "pages/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None                      # BASELINE.md holds no published number for this metric
    assert set(d["config"]) == {"workload", "l2", "parallelism"} and "Gundam" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "pages/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0
