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


# ---- control flow of the GPU arm with a stand-in device (no arithmetic): the line's keys, the guards, the ordering ----------
_FAKE_DRIVER = r'''
import json, os, sys, types, time
import numpy as np
sys.argv = ["bench.py"] + json.loads(os.environ["FAKE_ARGV"])
import bench
import torch
torch.cuda.synchronize = lambda *a, **k: None
import networksolvers_b200 as ns

FAIL = os.environ.get("FAKE_FAIL", "")

class FakeLib:
    def nsb_matvec_host(self, h, i, o): return 0
    nsb_matvec_host_slab = nsb_matvec_host

class FakeCtx:
    _lib = FakeLib()
    def __init__(self, dev=0): self.t = {}
    def set_option(self, k, v): pass
    def synchronize(self): pass
    def reset_counters(self): pass
    def gemm_profile(self, on): pass
    def profiler(self, on): pass
    def tic(self): pass
    def toc(self): return 12.0
    def gemm_profile_read(self): return [(0.5, 1e9, (64, 64, 64, 1)), (0.25, 5e8, (64, 32, 64, 1))] * 3
    def counters(self): return {"kernel_launches": 15, "gemm_calls": 9}
    def check(self, rc): assert rc == 0
    def dmma_peak_tflops(self): return 37.0
    def mem_info(self): return {"pool_used": 2**30}
    def enable_timers(self, on): pass
    def reset_timers(self): pass
    def timers(self): return {"matvec": 1.0, "factorize": 2.0}

class FakeNet:
    handle = None
    def __init__(self, chi): self.chi = chi
    def extract(self, region, *a):
        if FAIL == "region" and getattr(self, "_armed", False): raise RuntimeError("injected region failure")
        return types.SimpleNamespace(env_builds=3)
    def local_info(self): return ["l", "s", "s", "r"], [self.chi, 2, 2, self.chi]
    def matvec_flops(self): return bench.matvec_flops(self.chi, self.chi)
    def matvec_flops_executed(self): return 0.8 * self.matvec_flops()
    def matvec_device(self, reps, download=False): self._armed = True
    def shard_range(self): return 0, self.chi, self.chi
    def local_download(self): return np.zeros((self.chi, 2, 2, self.chi), order="F"), None
    def matvec_host(self, th): return th
    def update_eigsolve(self): return -1.0, None
    def insert(self, tr): return types.SimpleNamespace(newdim=self.chi)
    def env_bytes(self): return 10, 20
    def maxlinkdim(self): return self.chi

ns.Context = FakeCtx
bench.build_problem = lambda chi, nsites, ctx, **k: (FakeNet(chi), [nsites // 2, nsites // 2 + 1])
def fake_dmrg(prob, **kw):
    if FAIL == "sweep": raise RuntimeError("injected sweep failure")
    if FAIL == "hang": time.sleep(60)
    return -3.0, None
ns.dmrg = fake_dmrg
ns.EigsolveProblem = lambda net=None: types.SimpleNamespace(net=net)
bench.cpu_matvec_sample = lambda chi, seconds=12.0: (0.5, 0.01, 4, None)
bench.cpu_factorize_sample = lambda n, cutoff=0.0: 0.02
bench.main()
'''


def _fake_bench(argv, fail=""):
    env = dict(os.environ, FAKE_ARGV=json.dumps(argv), FAKE_FAIL=fail)
    r = subprocess.run([sys.executable, "-c", _FAKE_DRIVER], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1, r.stdout
    return lines[0]


CONTRACT_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                 "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline")


def test_gpu_arm_line_has_every_contract_key_and_the_measured_sweep():
    ln = _fake_bench(["--steps", "4", "--warmup", "1", "--chi", "8", "--nsites", "10"])
    for k in CONTRACT_KEYS:
        assert k in ln, k
    assert ln["warmup"] == 3                                  # W >= 3 enforced
    assert ln["steps"] == 4 and ln["n_gpus"] == 1 and ln["dtype"] == "f64" and ln["vs_baseline"] is None
    assert ln["ms_per_step"] == 3.0 and ln["gpu_launches"] == 15
    assert set(ln["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert ln["e2e"]["h2d_bytes_per_step"] == 8 * 2 * 2 * 8 * 8
    rf = ln["roofline"]
    assert rf["bound"] == "tensor" and rf["unit"] == "TFLOP/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-15
    assert abs(rf["achieved"] - (3 * 1.5e9) / (3 * 0.75) * 1e-9) < 1e-12        # issued flops / summed launch durations
    assert ln["cpu_baseline"]["kind"] == "port" and "region_step" in ln["cpu_baseline"]
    assert ln["full_sweep_regions"] == 18 and ln["full_sweep_energy"] == -3.0 and "full_sweep_error" not in ln
    assert len(ln["region_steps_s"]) == 6 and ln["sweep_regions"] == 18
    assert "workload" in ln["config"] and "model" not in ln["config"]


@pytest.mark.parametrize("fail,key", [("sweep", "full_sweep_error"), ("region", "region_step_error")])
def test_a_failing_optional_section_does_not_lose_the_headline_line(fail, key):
    ln = _fake_bench(["--steps", "2", "--warmup", "3", "--chi", "8", "--nsites", "10"], fail=fail)
    assert "injected" in ln[key]
    assert ln["value"] > 0 and ln["roofline"] is not None and "cpu_baseline" in ln


def test_watchdog_fires_when_the_sweep_hangs():
    ln = _fake_bench(["--steps", "2", "--warmup", "3", "--chi", "8", "--nsites", "10", "--sweep-watchdog-s", "1"], fail="hang")
    assert ln["full_sweep_error"].startswith("watchdog") and ln["value"] > 0 and "cpu_baseline" in ln
