"""bench.py contract that can be checked without a GPU: the reference arm (the CPU port of the
oracle) prints ONE JSON line with the agreed keys; under a multi-rank launch only rank 0 works."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
        "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def run(env_extra=None, *args):
    env = dict(os.environ)
    env.update(env_extra or {})
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-n", "32", "--pcg-iters", "10", *args],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    return [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    lines = run({"OMP_NUM_THREADS": "1"})          # what torchrun exports to its ranks
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert KEYS <= set(d)
    assert d["impl"] == "reference" and d["metric"] == "advect+project cell-updates/s"
    assert d["unit"] == "cell-updates/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    # all host cores although OMP_NUM_THREADS=1 was exported
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))


def test_reference_arm_other_ranks_exit_quietly():
    assert run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2") == []
