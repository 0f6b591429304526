"""CPU tests of bench.py's contract: the reference arm (oracle restatement on the host cores), its behaviour under a
torchrun-style environment, the watchdog line, and that the GPU arm has no CPU fallback."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def _run(argv, env_extra=None, timeout=300):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, BENCH] + argv, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


def _json_lines(out):
    return [json.loads(l) for l in out.splitlines() if l.startswith("{")]


def test_reference_arm_prints_one_contract_line_with_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm must still use every host core (N >= 2 ratios were void)."""
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1", "--chi", "256"],
             {"OMP_NUM_THREADS": "1", "RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1
    ln = lines[0]
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in ln, k
    assert ln["impl"] == "reference" and ln["metric"] == "heff_matvec_fp64_tflops" and ln["unit"] == "TFLOP/s"
    assert ln["n_gpus"] == 2 and ln["steps"] == 2 and ln["higher_is_better"] is True and ln["vs_baseline"] is None
    assert ln["config"]["chi"] == 256 and "workload" in ln["config"]
    assert ln["e2e"] == {"value": ln["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = ln["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == ln["value"] and "sample" in cb
    ncores = len(os.sched_getaffinity(0))
    assert cb["cores"] == ncores or ncores == 1, (cb["cores"], ncores)
    assert ln["value"] > 0 and ln["ms_per_step"] > 0


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--chi", "64"],
             {"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and _json_lines(r.stdout) == []


def test_watchdog_prints_the_line_without_the_sweep():
    code = ("import bench, sys\n"
            "bench._watchdog_emit(lambda extra, cpu: dict({'metric': 'm', 'cpu_baseline': cpu}, **extra), {'region_step_s': 1.0}, {'value': 2.0}, 5.0)\n"
            "print('not reached')\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120, cwd=ROOT)
    assert r.returncode == 0
    lines = _json_lines(r.stdout)
    assert len(lines) == 1 and "not reached" not in r.stdout
    assert lines[0]["metric"] == "m" and lines[0]["region_step_s"] == 1.0 and lines[0]["cpu_baseline"] == {"value": 2.0}
    assert lines[0]["full_sweep_error"].startswith("watchdog")


def test_gpu_arm_fails_loudly_without_a_device():
    """No CPU route behind the GPU arm: without a CUDA device there is no JSON line and a non-zero exit."""
    from helpers import cuda_available
    if cuda_available():
        pytest.skip("GPU present")
    r = _run(["--steps", "1", "--warmup", "1", "--chi", "16", "--nsites", "6", "--no-full-sweep", "--no-cpu-baseline"])
    assert r.returncode != 0
    assert _json_lines(r.stdout) == []
    assert "CUDA" in r.stderr or "cuda" in r.stderr
