"""CPU: the reference arm of bench.py (`--impl reference`: the oracle's OpenMP step + OpenMP LSTM act on the host cores) prints one JSON
line with the contract's keys; under torchrun only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None, args=()):
    env = dict(os.environ); env.update(env_extra or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1", "--cpu-envs", "64", *args],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip()


def test_reference_arm_line_has_the_contract_keys():
    out = _run()
    line = json.loads(out.splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "env-steps/sec incl. LSTM act" and line["unit"] == "env-steps/s"
    assert line["higher_is_better"] is True and line["steps"] == 3 and line["value"] > 0 and line["dtype"] == "f64"
    assert "configs[2]" in line["config"]["workload"] and line["config"]["envs"] == 64
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == line["value"] and cb["cores"] >= 1 and 1 <= cb["threads"] <= cb["cores"] and "64 envs" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_print_nothing():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == ""


def test_reference_arm_is_insensitive_to_omp_num_threads():
    """torchrun exports OMP_NUM_THREADS=1; the arm pins its own thread counts, so its thread sweep is the same with and without it"""
    a = json.loads(_run({"OMP_NUM_THREADS": "1"}).splitlines()[-1]); b = json.loads(_run().splitlines()[-1])
    assert a["cpu_baseline"]["threads"] in (a["cpu_baseline"]["cores"], max(1, a["cpu_baseline"]["cores"] // 2))
    assert a["cpu_baseline"]["cores"] == b["cpu_baseline"]["cores"]
