#!/usr/bin/env python
"""bench.py -- benchmarks of the sweep hot path on B200, one JSON line per run.

Default (BASELINE.json headline, config 2): S=1/2 Heisenberg chain N = 100, no QN, real FP64, chi = 4096.
A "step" is one projected effective-Hamiltonian application theta' = H_eff theta (L . W . W . R contraction,
src/operator_map.jl:3-10 of the reference) on an interior bond.  `value` = flops actually issued x steps / device time
with all operands resident in HBM; `e2e` = the same application through the reference-facing call with HOST buffers
(nsb_matvec_host: H2D copy of theta, matvec, D2H copy of theta') inside the timed region.  The same run then times
consecutive full region steps (extract -> 3-matvec Lanczos -> truncating insert) and, on one GPU, one complete measured
2-site DMRG sweep over all 2 (N - 1) regions.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference          # the reference-equivalent CPU path on the host cores (all threads)
    python bench.py --config 1|3|4|5          # the other BASELINE configs (own metric each, same JSON contract)

`roofline` is derived from the GEMM launches of the timed steps themselves: every launch is bracketed by CUDA events on
the launching stream (nsb_gemm_profile_*), achieved = sum of issued flops / sum of launch durations."""
import os
import sys


def _host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


if "reference" in sys.argv:
    # torchrun exports OMP_NUM_THREADS=1: the CPU arm must use every host core whatever the launcher set (read at BLAS load)
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = str(_host_cores())

import argparse
import json
import subprocess
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W_MPO, D_SITE = 5, 2
CPU_SLABS = 8     # the CPU sample is one right-bond slab (1 / CPU_SLABS) of the matvec at the GPU arm's chi


def matvec_flops(chi_l, chi_r, d=D_SITE, w=W_MPO, cplx=False):
    """Algorithmic flops of the fixed-order matvec for theta[chi_l, d, d, chi_r]."""
    f = 2.0 * w * chi_l * chi_l * d * d * chi_r          # L . theta
    f += 2.0 * chi_l * d * chi_r * (w * d) * (d * w)       # W1
    f += 2.0 * chi_l * d * chi_r * (w * d) * (d * w)       # W2
    f += 2.0 * chi_l * d * d * chi_r * w * chi_r           # . R
    return f * (4.0 if cplx else 1.0)


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nm, val in zip(names, r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement of the reference path on the host cores
# ------------------------------------------------------------------------------------------------------------------
def force_blas_threads():
    """All host cores for the BLAS behind NumPy, whatever OMP_NUM_THREADS said when it was loaded."""
    n = _host_cores()
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n)
    except Exception:
        pass
    return blas_threads()


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [p.get("num_threads") for p in threadpool_info() if p.get("user_api") == "blas"]
        return max(n) if n else _host_cores()
    except Exception:
        return _host_cores()


def blas_name():
    try:
        from threadpoolctl import threadpool_info
        for p in threadpool_info():
            if p.get("user_api") == "blas":
                return f"{p.get('internal_api')} {p.get('version')}"
    except Exception:
        pass
    return "unknown"


class CpuSlabMatvec:
    """Bounded CPU sample of the chi-sized matvec: one right-bond slab (1 / nslabs of theta's last bond, what one rank of the
    sharded GPU application computes) through the oracle's optimal_map (src/operator_map.jl:3-10 restated) with synthetic
    environments of the full bond dimension.  Flops = dense matvec count / nslabs exactly."""

    def __init__(self, chi, nslabs=CPU_SLABS, cplx=False, w=W_MPO, d=D_SITE):
        from oracle.projttn import ProjTTN
        from oracle.models import TTN
        from oracle.graph import path_graph
        from oracle.tensor import Tensor, site, link, oplink
        rng = np.random.default_rng(1234)
        self.chi, self.nslabs, self.cplx = chi, nslabs, cplx
        sl = max(chi // nslabs, 1)

        def rnd(shape):
            a = rng.standard_normal(shape)
            return a + 1j * rng.standard_normal(shape) if cplx else a

        g = path_graph(4)
        Wt = {v: Tensor(rnd((w, w, d, d)), [oplink(v - 1, v), oplink(v, v + 1), site(v, 0), site(v, 1)]) for v in (2, 3)}
        Hn = TTN(g, {1: None, 2: Wt[2], 3: Wt[3], 4: None}, ortho_region=[])
        self.P = ProjTTN(Hn, pos=[2, 3])
        self.P.environments[(1, 2)] = Tensor(rnd((chi, w, chi)) / chi, [link(1, 2, 0), oplink(1, 2), link(1, 2, 1)])
        self.P.environments[(4, 3)] = Tensor(rnd((sl, w, chi)) / chi, [link(3, 4, 0), oplink(3, 4), link(3, 4, 1)])
        self.theta = Tensor(rnd((chi, d, d, sl)), [link(1, 2), site(2), site(3), link(3, 4)])
        self.flops = matvec_flops(chi, chi, d, w, cplx) * sl / chi

    def step(self):
        from oracle.operator_map import optimal_map
        t0 = time.perf_counter()
        optimal_map(self.P, self.theta)
        return time.perf_counter() - t0


def cpu_matvec_sample(chi, seconds=12.0, min_reps=2, cplx=False):
    force_blas_threads()
    s = CpuSlabMatvec(chi, cplx=cplx)
    s.step()                                   # warm-up (BLAS thread start-up, page faults)
    ts = []
    t_end = time.perf_counter() + seconds
    while len(ts) < min_reps or time.perf_counter() < t_end:
        ts.append(s.step())
        if len(ts) >= 50:
            break
    dt = float(np.mean(ts))
    return s.flops / dt * 1e-12, dt, len(ts), s


def cpu_factorize_sample(n, cutoff=0.0):
    """Oracle `factorize` (LAPACK SVD for cutoff <= 1e-12, density-matrix eigen above; ITensors rule App. A.4) of an n x n
    two-site tensor, maxdim n / 2."""
    from oracle.tensor import Tensor, factorize, link, site
    force_blas_threads()
    rng = np.random.default_rng(4321)
    chi = n // D_SITE
    theta = Tensor(rng.standard_normal((chi, D_SITE, D_SITE, chi)), [link(1, 2), site(2), site(3), link(3, 4)])
    t0 = time.perf_counter()
    factorize(theta, [link(1, 2), site(2)], link(2, 3), cutoff=cutoff, maxdim=chi)
    return time.perf_counter() - t0


def run_reference(args, rank, world):
    """`--impl reference`: the reference's own implementation cannot run here (Julia + un-vendored ITensors / KrylovKit), so
    this times the oracle restatement (kind = "port") on every host core, rank 0 only.  Same config as the GPU arm (chi);
    each step is a bounded sample of it: one 1/8 right-bond slab of the matvec (flops counted accordingly)."""
    if rank != 0:
        return
    cores = force_blas_threads()
    chi = args.chi
    s = CpuSlabMatvec(chi)
    for _ in range(max(args.warmup, 1)):
        s.step()
    ts = [s.step() for _ in range(max(args.steps, 1))]
    dt = float(np.mean(ts))
    tf = s.flops / dt * 1e-12
    sample = (f"{len(ts)} steps, each one right-bond slab (1/{CPU_SLABS} of theta, {s.flops:.3e} flop) of the chi={chi} H_eff matvec "
              f"(d=2, w=5) through the oracle restatement of optimal_map, NumPy + {blas_name()}, {cores} threads of {_host_cores()} cores")
    line = {"impl": "reference", "metric": "heff_matvec_fp64_tflops", "value": tf, "unit": "TFLOP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "chi": chi,
                       "sample": f"1/{CPU_SLABS} right-bond slab of the matvec per step; rate metric"},
            "cpu_baseline": {"value": tf, "unit": "TFLOP/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": tf, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def workload_name(args):
    return (f"S=1/2 Heisenberg chain N={args.nsites}, no QN, 2-site H_eff matvec on an interior bond, chi={args.chi}, d=2, w=5 "
            f"(BASELINE config 2)")


# ------------------------------------------------------------------------------------------------------------------
# helpers of the GPU arm
# ------------------------------------------------------------------------------------------------------------------
def pinned_array(shape, dtype):
    try:
        import torch
        t = torch.empty(int(np.prod(shape)), dtype=torch.float64 if np.dtype(dtype).kind == "f" else torch.complex128,
                        pin_memory=True)
        return t.numpy().reshape(shape, order="F"), t
    except Exception:
        return np.empty(shape, dtype=dtype, order="F"), None


def emit(line):
    """The one JSON line goes to the process's original stdout; everything else (NCCL's version banner, library
    chatter written to fd 1 from C) was redirected to stderr at start-up."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def _watchdog_emit(make_line, extra, cpu, seconds):
    """The full sweep did not return within `seconds`: print the line of what was measured and leave."""
    extra = dict(extra)
    extra["full_sweep_error"] = f"watchdog: sweep not finished after {seconds:.0f} s, line printed without it"
    try:
        emit(make_line(extra, cpu))
    finally:
        os._exit(0)


_REAL_STDOUT = 1


def roofline_from_profile(ctx, recs, ms_total, traffic_key=None):
    """Roofline of the dominant kernel (the persistent TMA + DMMA GEMM) from the launches of the timed region."""
    if not recs:
        return None
    peak = ctx.dmma_peak_tflops()
    g_ms = sum(r[0] for r in recs)
    g_fl = sum(r[1] for r in recs)
    by = {}
    for ms, fl, mnk in recs:
        e = by.setdefault(mnk, [0, 0.0, 0.0])
        e[0] += 1; e[1] += ms; e[2] += fl
    launches = [{"M": k[0], "N": k[1], "K": k[2], "batch": k[3], "launches": v[0], "avg_ms": v[1] / v[0],
                 "tflops": v[2] / v[1] * 1e-9 if v[1] > 0 else None} for k, v in sorted(by.items(), key=lambda kv: -kv[1][1])]
    achieved = g_fl / g_ms * 1e-9
    traffic = None
    tnote = "no ncu --set full capture of this launch committed for this round"
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))
        if traffic_key and traffic_key in tj:
            traffic = tj[traffic_key]["dram_bytes"]
            tnote = tj[traffic_key].get("note", "")
    except Exception:
        pass
    return {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_note": tnote,
            "kernel": "gemm_tma_kernel (persistent TMA + mbarrier + DMMA GEMM): every GEMM launch of the timed steps (launches below m n k = 2e7 -- only the small-chi configs have them -- run on the plain FP64 kernel)",
            "how": "CUDA events on the launching stream around each GEMM launch inside the timed region; achieved = issued flops / "
                   "summed launch durations",
            "gemm_ms_per_step_share": g_ms / ms_total if ms_total > 0 else None,
            "launch_shapes": launches[:6],
            "peak_source": "FP64 DMMA issue ceiling measured live by nsb_dmma_peak (register-resident mma.sync loop); "
                           "MEASURED_PEAKS.json has no FP64 entry and the profiling guide states no FP64 fallback "
                           "(148 SM x 64 FMA x 2 x 1.965 GHz = 37.2)"}


def build_problem(chi, nsites, ctx, dtype=np.float64, canonical=True, model="heisenberg"):
    import networksolvers_b200 as ns
    g = ns.path_graph(nsites)
    sites = ns.siteinds("S=1/2", g)
    H = ns.ttno(ns.heisenberg(g) if model == "heisenberg" else ns.transverse_ising(g, 1.0, 1.0), sites)
    mid = nsites // 2
    region = [mid, mid + 1]
    net = ns.DeviceNetwork.synthetic(H, sites, chi, seed=1234, dtype=dtype, ctx=ctx, ortho_region=region, canonical=canonical)
    return net, region


def time_region_steps(ctx, net, regions, trunc, solver):
    """Full region steps through the three hooks; returns (wall seconds per step, phase timers per step, newdims)."""
    ctx.enable_timers(True)
    steps, phases, newdims = [], [], []
    marked = os.environ.get("NSB_PROFILE_REGION") == "1"
    for ir, reg in enumerate(regions):
        if marked and ir == 2:
            ctx.profiler(True)               # the third step: a steady-state region step with one environment update
        ctx.reset_timers()
        ctx.synchronize()
        t0 = time.perf_counter()
        net.extract(reg)
        solver()
        ins = net.insert(trunc)
        ctx.synchronize()
        steps.append(time.perf_counter() - t0)
        phases.append(ctx.timers())
        newdims.append(int(ins.newdim))
        if marked and ir == 2:
            ctx.profiler(False)
    ctx.enable_timers(False)
    return steps, phases, newdims


# ------------------------------------------------------------------------------------------------------------------
# config 2 (headline)
# ------------------------------------------------------------------------------------------------------------------
def run_config2(args, rank, world, local_rank):
    import torch
    import networksolvers_b200 as ns
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = ns.Context(local_rank)
    for kv in args.opt:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    t_setup = time.perf_counter()
    net, region = build_problem(args.chi, args.nsites, ctx, canonical=not args.no_canonical)   # same seed on every rank
    info = net.extract(region)
    ctx.synchronize()
    t_setup = time.perf_counter() - t_setup
    legs, dims = net.local_info()
    flops_dense = net.matvec_flops()
    assert abs(flops_dense - matvec_flops(dims[0], dims[-1])) < 1e-6 * flops_dense, (flops_dense, dims)
    shard = None
    shard_err = None
    shard_err_repeat = None
    if world > 1:
        from networksolvers_b200.parallel import setup_sharded_matvec
        shard = setup_sharded_matvec(net, dist, rank, world, fused=args.fused)
        if not shard.active:
            shard = None
        else:
            # driver-side multi-GPU parity: sharded application against the single-GPU application of the same theta
            y_sh = net.matvec_device(1, download=True)
            shard.enable(False)
            y_rep = net.matvec_device(1, download=True)
            shard.enable(True)
            err_local = float(np.abs(y_sh - y_rep).max() / np.abs(y_rep).max())
            t = torch.tensor([err_local], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            shard_err = float(t.item())
            if shard_err > 1e-12:      # (decision identical on every rank) evidence for the log: persistent or transient?
                again = [float(np.abs(net.matvec_device(1, download=True) - y_rep).max() / np.abs(y_rep).max()) for _ in range(2)]
                dlt = np.abs(y_sh - y_rep)
                per = dlt.shape[-1] // world
                print(f"[rank {rank}] sharded-vs-single error {err_local:.3e} (max over ranks {shard_err:.3e}); repeated applications: {again}; "
                      f"per-slab max abs err {[float(dlt[..., r * per:(r + 1) * per].max()) for r in range(world)]}, max|y| {float(np.abs(y_rep).max()):.3e}",
                      file=sys.stderr, flush=True)
                t = torch.tensor([min(again)], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                shard_err_repeat = float(t.item())
                if os.environ.get("NSB_BENCH_STRICT") == "1":
                    assert shard_err_repeat <= 1e-12, f"sharded matvec differs from the single-GPU one: {shard_err} (repeats {again})"
            del y_sh, y_rep

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.synchronize()

    def step():
        net.matvec_device(1)

    for _ in range(args.warmup):
        step()
    barrier()
    flops = net.matvec_flops_executed()   # what the applications above really issued (whole job, all ranks)
    ctx.reset_counters()
    ctx.gemm_profile(True)
    marked = os.environ.get("NSB_PROFILE_TIMED") == "1"      # ncu --profile-from-start off: only the timed steps are profiled
    if marked:
        ctx.profiler(True)
    with ClockSampler(local_rank) as clk:
        ctx.tic()
        for _ in range(args.steps):
            step()
        ms_total = ctx.toc()
    if marked:
        ctx.profiler(False)
    barrier()
    recs = ctx.gemm_profile_read()
    ctx.gemm_profile(False)
    launches = ctx.counters()["kernel_launches"]
    ms_local = ms_total
    if dist is not None:
        tmax = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms_total = float(tmax.item())
    ms_step = ms_total / args.steps
    tflops = flops / (ms_step * 1e-3) * 1e-12
    replicas = world if (shard is None and world > 1) else 1

    # ---- end-to-end through the host-buffer call ----
    # Multi-GPU: the caller holds theta distributed like the sharded Krylov solvers do -- every rank passes its slab of the last
    # bond through nsb_matvec_host_slab (1 / N of the PCIe traffic per rank); single GPU: the whole vector (nsb_matvec_host).
    lo, hi, last_dim = net.shard_range()
    sdims = list(dims[:-1]) + [hi - lo]
    nbytes = int(np.prod(sdims)) * 8            # per rank
    hin, keep1 = pinned_array(sdims, np.float64)
    theta0, _ = net.local_download()
    hin[...] = theta0[..., lo:hi]
    hout, keep2 = pinned_array(sdims, np.float64)
    lib = ctx._lib
    e2e_call = lib.nsb_matvec_host_slab if shard is not None else lib.nsb_matvec_host
    e2e_steps = max(3, min(args.steps, 5))
    for _ in range(2):
        ctx.check(e2e_call(net.handle, hin.ctypes.data, hout.ctypes.data))
    if shard is not None and rank == 0:          # the slab call against the replicated call on the same input
        yfull = net.matvec_host(theta0)
        slab_err = float(np.abs(hout - yfull[..., lo:hi]).max() / np.abs(yfull).max())
        assert slab_err <= 1e-12 or os.environ.get("NSB_BENCH_STRICT") != "1", slab_err
    elif shard is not None:
        net.matvec_host(theta0)                  # (collective inside: every rank takes part)
        slab_err = None
    else:
        slab_err = None
    del theta0
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.check(e2e_call(net.handle, hin.ctypes.data, hout.ctypes.data))
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if dist is not None:
        tt = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e_tflops = flops * replicas / e2e_s * 1e-12

    roof = roofline_from_profile(ctx, recs, ms_local, traffic_key=f"chi{args.chi}_matvec_gemm") if rank == 0 else None

    extra = {}
    if not args.no_region_step and (world == 1 or shard is not None):
        # (guarded: whatever happens below, the headline line of the timed matvec steps above is still printed)
        try:
            # Consecutive full region steps of a 2-site sweep through the three hooks, first to the right from the benchmark
            # bond, then back to the left (both directions of the Euler tour).
            tr = (args.cutoff, 1, args.chi)
            nr = max(args.region_steps, 1)
            right = [[region[0] + r, region[1] + r] for r in range(nr) if region[1] + r <= args.nsites]
            left = [[right[-1][1] - r, right[-1][0] - r] for r in range(nr)]
            steps, phases, newdims = time_region_steps(ctx, net, right + left, tr, lambda: net.update_eigsolve())
            full = list(range(1, len(right))) + list(range(len(right) + 1, len(steps)))   # steps that include one environment update
            # medians: the first step that updates an environment after set-up also pays one-off allocations of the step's work buffers
            extra["region_step_s"] = float(np.median([steps[i] for i in full]))
            extra["region_steps_s"] = steps
            extra["region_directions"] = ["right"] * len(right) + ["left"] * len(left)
            extra["region_phase_ms"] = {k: float(np.median([phases[i][k] for i in full])) for k in phases[0]}
            extra["region_newdim"] = newdims
            extra["region_trunc"] = {"cutoff": args.cutoff, "maxdim": args.chi}
            extra["sweep_regions"] = 2 * (args.nsites - 1)
            extra["sweep_s_extrapolated"] = extra["region_step_s"] * 2 * (args.nsites - 1)
            if world > 1:
                extra["region_parallelism"] = (
                    "Krylov vectors sharded along theta's last bond (ncclReduceScatter when sweeping right, ncclAllGather when sweeping left, "
                    "scalar all-reduce per dot); environment update split over the incoming environment's bra index + all-reduce; "
                    "factorisation: Gram matrix, back-transformation and C = U^H theta by column slabs + all-gather, tridiagonalisation "
                    "and divide & conquer replicated; tensors replicated in HBM")
            mi = ctx.mem_info()
            extra["hbm_pool_used_gib"] = mi["pool_used"] / 2**30
            er, ef = net.env_bytes()
            extra["env_hbm_gib_per_gpu"] = er / 2**30
            extra["env_hbm_gib_if_replicated"] = ef / 2**30
        except Exception as e:      # noqa: BLE001
            if world > 1:
                raise                # (ranks in lock step: one rank carrying on alone would leave the others inside a collective)
            extra["region_step_error"] = f"{type(e).__name__}: {e}"[:300]

    def make_line(extra, cpu):
        line = {"metric": "heff_matvec_fp64_tflops", "value": tflops * replicas,
                "unit": "TFLOP/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong" if shard is not None else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(args), "bond": region,
                           "chi": args.chi, "local_dims": dims, "flops_per_step": flops, "dense_flops_per_step": flops_dense,
                           "flops_note": "value / e2e use the flops actually issued; dense_flops_per_step is the reference's "
                                         "dense-W count 4 w d^2 chi^3 + 4 w^2 d^3 chi^2 (identity channels of L and R skipped)",
                           "l2": "inputs larger than L2 (L 0.64 GB, theta 0.5 GB, T1 2.5 GB per matvec)",
                           "state": ("random tensors orthonormalised to the benchmark bond (device QR gauge walk)" if not args.no_canonical else "random tensors, gauge flag only"),
                           "parallelism": ("replicated" if shard is None else net_parallelism(net, world, args))
                           if world > 1 else "single GPU", "setup_s": t_setup, "env_builds": info.env_builds},
                "e2e": {"value": e2e_tflops, "unit": "TFLOP/s", "h2d_bytes_per_step": nbytes * (world if shard is not None else 1),
                        "d2h_bytes_per_step": nbytes * (world if shard is not None else 1), "ms_per_step": e2e_s * 1e3,
                        "call": ("nsb_matvec_host_slab: every rank moves its slab of theta / theta' (bytes are the whole job's, "
                                 f"{nbytes} per rank)" if shard is not None else "nsb_matvec_host"),
                        **({"slab_vs_full_call_max_rel_err": slab_err} if slab_err is not None else {})},
                "gpu_launches": int(launches), "clocks": clk.summary(), "roofline": roof,
                "dense_equivalent_tflops": flops_dense / (ms_step * 1e-3) * 1e-12 * replicas}
        if shard_err is not None:
            line["sharded_vs_replicated_max_rel_err"] = shard_err
            line["sharded_parity_ok"] = bool(shard_err <= 1e-12)
            if shard_err_repeat is not None:
                line["sharded_vs_replicated_repeat_err"] = shard_err_repeat
        if cpu is not None:
            line["cpu_baseline"] = cpu
        line.update(extra)
        return line

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            ctf, cdt, creps, s = cpu_matvec_sample(args.chi, seconds=12.0)
            cores = blas_threads()
            cpu = {"value": ctf, "unit": "TFLOP/s", "cores": cores, "kind": "port",
                   "sample": f"{creps} x one right-bond slab (1/{CPU_SLABS}) of the chi={args.chi} H_eff matvec, oracle restatement of "
                             f"optimal_map (NumPy + {blas_name()}, {cores} threads); Julia reference not runnable here"}
            if not args.no_region_step:
                # factorisation sample at a bounded size (LAPACK gesdd / syevd scale as n^3), scaled to the GPU arm's 2 chi
                nf = min(2 * args.chi, 2048)
                dt_f = cpu_factorize_sample(nf, cutoff=args.cutoff)
                mv_full = cdt * CPU_SLABS
                cpu["region_step"] = {"matvec_s_at_chi": mv_full, "factorize_s_sample": dt_f, "factorize_sample_n": nf,
                                      "factorize_s_scaled": dt_f * (2 * args.chi / nf) ** 3,
                                      "region_s_estimate": 3 * mv_full + dt_f * (2 * args.chi / nf) ** 3,
                                      "note": "3 matvecs (slab sample x 8) + oracle factorize (LAPACK) scaled by (n / n_sample)^3; "
                                              "no environment update counted"}
        except Exception as e:      # noqa: BLE001
            cpu = None
            extra["cpu_baseline_error"] = f"{type(e).__name__}: {e}"[:300]

    if not args.no_full_sweep and (world == 1 or args.full_sweep):
        # A watchdog prints the line without the sweep if the sweep does not come back (the matvec numbers are the headline).
        dog = threading.Timer(args.sweep_watchdog_s, _watchdog_emit, args=(make_line, extra, cpu, args.sweep_watchdog_s))
        dog.daemon = True
        if rank == 0:
            dog.start()
        try:
            # One real 2-site DMRG sweep (all 2 (N - 1) regions of the Euler tour) through the public driver, continuing on the
            # same network: the gauge walk to the tour's first region and the environments it needs are set-up (not timed).
            g = ns.path_graph(args.nsites)
            plan = ns.euler_sweep(g, nsites=2)
            t0 = time.perf_counter()
            net.extract(list(plan[0][0]))
            ctx.synchronize()
            extra["full_sweep_setup_s"] = time.perf_counter() - t0
            prob = ns.EigsolveProblem(net=net)
            ctx.enable_timers(True)
            ctx.reset_timers()
            ctx.reset_counters()
            barrier()
            t0 = time.perf_counter()
            E, _ = ns.dmrg(prob, nsweeps=1, nsites=2, inserter_kwargs=dict(trunc=dict(cutoff=args.cutoff, maxdim=args.chi)))
            barrier()
            extra["full_sweep_s"] = time.perf_counter() - t0
            extra["full_sweep_regions"] = len(plan)
            extra["full_sweep_phase_ms"] = ctx.timers()
            extra["full_sweep_launches"] = int(ctx.counters()["kernel_launches"])
            extra["full_sweep_maxlinkdim"] = int(net.maxlinkdim())
            extra["full_sweep_energy"] = float(E)
            ctx.enable_timers(False)
        except Exception as e:      # noqa: BLE001
            if world > 1:
                raise
            extra["full_sweep_error"] = f"{type(e).__name__}: {e}"[:300]
        finally:
            dog.cancel()

    if rank == 0:
        emit(make_line(extra, cpu))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def net_parallelism(net, world, args):
    if args.fused:
        return f"theta right-bond sharded x{world} + fused GEMM/peer-store reduce-scatter + allgather"
    return f"theta last-bond sharded x{world} + NCCL collective per application (see region_parallelism)"


# ------------------------------------------------------------------------------------------------------------------
# other BASELINE configs (single GPU)
# ------------------------------------------------------------------------------------------------------------------
def _line(metric, value, unit, args, ms_step, hib, dtype, workload, extra_cfg, e2e, launches, clk, roof, cpu, extra):
    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": hib, "scaling": "weak", "vs_baseline": None, "dtype": dtype,
            "data": "synthetic", "config": dict({"workload": workload}, **extra_cfg), "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clk, "roofline": roof}
    if cpu is not None:
        line["cpu_baseline"] = cpu
    line.update(extra)
    return line


def run_config1(args):
    """BASELINE config 1: S=1/2 Heisenberg N=20, 2-site DMRG, maxdim 100, cutoff 1e-12 from the Neel product state (the
    reference's own CPU-runnable case).  A step = one full sweep at saturated bond dimension; e2e = the whole public call
    ns.dmrg(H, psi0) from host tensors (upload, 5 sweeps, download) per sweep."""
    import networksolvers_b200 as ns
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import neel, to_oracle_ttn
    ctx = ns.default_context()
    g = ns.path_graph(20)
    sites = ns.siteinds("S=1/2", g)
    H = ns.ttno(ns.heisenberg(g), sites)
    psi0 = ns.product_state(sites, neel(g))
    trunc = dict(cutoff=1e-12, maxdim=100)
    prob = ns.EigsolveProblem(state=psi0, operator=H, ctx=ctx)
    E, _ = ns.dmrg(prob, nsweeps=max(args.warmup, 4), nsites=2, inserter_kwargs=dict(trunc=trunc))
    ctx.reset_counters()
    ctx.gemm_profile(True)
    with ClockSampler(0) as clk:
        ctx.tic()
        t0 = time.perf_counter()
        E, _ = ns.dmrg(prob, nsweeps=args.steps, nsites=2, inserter_kwargs=dict(trunc=trunc))
        ctx.synchronize()
        wall = time.perf_counter() - t0
        ms_total = ctx.toc()
    recs = ctx.gemm_profile_read()
    ctx.gemm_profile(False)
    launches = ctx.counters()["kernel_launches"]
    roof = roofline_from_profile(ctx, recs, ms_total)
    if roof:
        roof["note"] = "chi <= 100: the sweep is launch-latency bound, the GEMM share of the step says so"
    t0 = time.perf_counter()
    E2, psi = ns.dmrg(H, psi0, nsweeps=5, nsites=2, inserter_kwargs=dict(trunc=trunc), ctx=ctx)
    host = psi.to_host()
    e2e_s = (time.perf_counter() - t0) / 5
    nbytes_in = sum(a.nbytes for a in psi0.tensors.values()) + sum(a.nbytes for a in H.tensors.values())
    nbytes_out = sum(a.nbytes for a in host.tensors.values())
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import sweep as osw
        cores = force_blas_threads()
        t0 = time.perf_counter()
        Eo, _ = osw.dmrg(to_oracle_ttn(H, True), to_oracle_ttn(psi0), nsweeps=5, nsites=2, inserter_kwargs=dict(trunc=trunc))
        cs = (time.perf_counter() - t0) / 5
        cpu = {"value": cs, "unit": "s/sweep", "cores": cores, "kind": "port",
               "sample": f"oracle dmrg, the same 5 sweeps from the Neel state (NumPy + {blas_name()}), energy {Eo:.12f}"}
    extra = {"energy": float(E2), "energy_exact_ed": -8.682473334399, "sweep_wall_s": wall / args.steps}
    emit(_line("dmrg_two_site_sweep_s", ms_total / args.steps * 1e-3, "s/sweep", args, ms_total / args.steps, False, "f64",
               "S=1/2 Heisenberg chain N=20 path_graph, 2-site DMRG maxdim=100 cutoff=1e-12 (BASELINE config 1, examples/dmrg.jl shape)",
               {"maxlinkdim": int(prob.net.maxlinkdim()), "l2": "working set far below L2: latency-bound regime, nothing to flush"},
               {"value": e2e_s, "unit": "s/sweep", "h2d_bytes_per_step": nbytes_in // 5, "d2h_bytes_per_step": nbytes_out // 5},
               launches, clk.summary(), roof, cpu, extra))


def run_config4(args):
    """BASELINE config 4: 2-site TDVP quench, N=64, chi=1024, complex128, Heisenberg (w=5); dt=0.05, RK4 local solver, cutoff
    1e-14 (examples/quench_evolution.jl:20-57).  A step = one complex H_eff application; then 2-site TDVP region steps
    (4 matvecs + truncating factorisation)."""
    import networksolvers_b200 as ns
    ctx = ns.default_context()
    chi, N = args.chi if args.chi != 4096 else 1024, 64
    out = {}
    for model in ("heisenberg", "ising"):
        w = 5 if model == "heisenberg" else 3
        t0 = time.perf_counter()
        net, region = build_problem(chi, N, ctx, dtype=np.complex128, canonical=not args.no_canonical, model=model)
        net.extract(region)
        ctx.synchronize()
        setup = time.perf_counter() - t0
        legs, dims = net.local_info()
        for _ in range(args.warmup):
            net.matvec_device(1)
        flops = net.matvec_flops_executed()
        ctx.reset_counters()
        ctx.gemm_profile(True)
        with ClockSampler(0) as clk:
            ctx.tic()
            net.matvec_device(args.steps)
            ms_total = ctx.toc()
        recs = ctx.gemm_profile_read()
        ctx.gemm_profile(False)
        launches = ctx.counters()["kernel_launches"]
        ms = ms_total / args.steps
        roof = roofline_from_profile(ctx, recs, ms_total)
        hin, k1 = pinned_array(dims, np.complex128)
        th, _ = net.local_download()
        hin[...] = th
        hout, k2 = pinned_array(dims, np.complex128)
        for _ in range(2):
            ctx.check(ctx._lib.nsb_matvec_host(net.handle, hin.ctypes.data, hout.ctypes.data))
        t0 = time.perf_counter()
        for _ in range(3):
            ctx.check(ctx._lib.nsb_matvec_host(net.handle, hin.ctypes.data, hout.ctypes.data))
        e2e_s = (time.perf_counter() - t0) / 3
        nbytes = int(np.prod(dims)) * 16
        tr = (1e-14, 1, chi)
        regs = [[region[0] + r, region[1] + r] for r in range(3)]
        steps, phases, newdims = time_region_steps(ctx, net, regs, tr, lambda: net.update_exp(-0.05j, solver="rk", order=4, nsites=2))
        res = {"matvec_ms": ms, "matvec_tflops_real": flops / ms * 1e-9, "setup_s": setup, "region_step_s": float(np.mean(steps[1:])),
               "region_phase_ms": {k: float(np.mean([p[k] for p in phases[1:]])) for k in phases[0]}, "region_newdim": newdims,
               "half_sweep_s_extrapolated": float(np.mean(steps[1:])) * (N - 1), "time_step_order4_s_extrapolated": float(np.mean(steps[1:])) * (N - 1) * 6}
        if model == "heisenberg":
            cpu = None
            if not args.no_cpu_baseline:
                ctf, cdt, creps, s = cpu_matvec_sample(chi, seconds=10.0, cplx=True)
                cpu = {"value": ctf, "unit": "TFLOP/s", "cores": blas_threads(), "kind": "port",
                       "sample": f"{creps} x one right-bond slab (1/{CPU_SLABS}) of the complex chi={chi} matvec, oracle optimal_map (NumPy + {blas_name()})"}
            head = dict(ms=ms, flops=flops, e2e_s=e2e_s, nbytes=nbytes, launches=launches, clk=clk.summary(), roof=roof, cpu=cpu, dims=dims)
        out[model] = res
        net.close()
    emit(_line("heff_matvec_c128_real_tflops", head["flops"] / head["ms"] * 1e-9, "TFLOP/s", args, head["ms"], True, "c128",
               f"2-site TDVP quench on the S=1/2 Heisenberg chain N=64, chi={chi}, complex128, dt=0.05, RK4 (BASELINE config 4, "
               "examples/quench_evolution.jl shape); Ising (w=3) beside it",
               {"chi": chi, "local_dims": head["dims"], "flops_per_step": head["flops"], "flops_note": "real flops (4 per complex multiply-add pair)",
                "l2": "inputs larger than L2 (T1 320 MiB per matvec)"},
               {"value": head["flops"] / head["e2e_s"] * 1e-12, "unit": "TFLOP/s", "h2d_bytes_per_step": head["nbytes"], "d2h_bytes_per_step": head["nbytes"],
                "ms_per_step": head["e2e_s"] * 1e3},
               head["launches"], head["clk"], head["roof"], head["cpu"], {"models": out}))


def run_config5(args):
    """BASELINE config 5: Heisenberg on named_comb_tree (10 teeth x 6 = 60 sites), Euler-tour plan.  As SURVEY 8(d) states the
    2-site tensor on a backbone edge at chi=512 is 2 TiB, so the run is 1-site DMRG + "densitymatrix" expansion at chi=512 on a
    degree-3 backbone vertex (local tensor chi^3 d = 2 GiB).  A step = one H_eff application on that vertex."""
    import networksolvers_b200 as ns
    ctx = ns.default_context()
    chi = args.chi if args.chi != 4096 else 512
    g = ns.named_comb_tree([6] * 10)
    sites = ns.siteinds("S=1/2", g)
    H = ns.ttno(ns.heisenberg(g), sites)
    v = (5, 1)
    t0 = time.perf_counter()
    net = ns.DeviceNetwork.synthetic(H, sites, chi, seed=7, dtype=np.float64, ctx=ctx, ortho_region=[v], canonical=not args.no_canonical)
    net.extract([v])
    ctx.synchronize()
    setup = time.perf_counter() - t0
    legs, dims = net.local_info()
    for _ in range(args.warmup):
        net.matvec_device(1)
    flops = net.matvec_flops_executed()
    ctx.reset_counters()
    ctx.gemm_profile(True)
    with ClockSampler(0) as clk:
        ctx.tic()
        net.matvec_device(args.steps)
        ms_total = ctx.toc()
    recs = ctx.gemm_profile_read()
    ctx.gemm_profile(False)
    c = ctx.counters()
    ms = ms_total / args.steps
    roof = roofline_from_profile(ctx, recs, ms_total)
    hin, k1 = pinned_array(dims, np.float64)
    th, _ = net.local_download()
    hin[...] = th
    hout, k2 = pinned_array(dims, np.float64)
    for _ in range(2):
        ctx.check(ctx._lib.nsb_matvec_host(net.handle, hin.ctypes.data, hout.ctypes.data))
    t0 = time.perf_counter()
    for _ in range(3):
        ctx.check(ctx._lib.nsb_matvec_host(net.handle, hin.ctypes.data, hout.ctypes.data))
    e2e_s = (time.perf_counter() - t0) / 3
    nbytes = int(np.prod(dims)) * 8
    ctx.enable_timers(True)
    ctx.reset_timers()
    t0 = time.perf_counter()
    val, sinfo = net.update_eigsolve()
    net.insert((1e-9, 1, chi))
    ctx.synchronize()
    region_s = time.perf_counter() - t0
    ph = ctx.timers()
    ctx.enable_timers(False)
    extra = {"region_step_s": region_s, "region_phase_ms": ph, "nmatvec": int(sinfo.nmatvec), "setup_s": setup,
             "permute_bytes_per_matvec": c["permute_bytes"] / args.steps, "hbm_pool_used_gib": ctx.mem_info()["pool_used"] / 2**30}
    emit(_line("heff_matvec_fp64_tflops", flops / ms * 1e-9, "TFLOP/s", args, ms, True, "f64",
               f"Heisenberg on named_comb_tree 10 x 6 (60 sites), 1-site H_eff on the degree-3 backbone vertex {v}, chi={chi} "
               "(BASELINE config 5; the 2-site tensor on a backbone edge would be chi^4 d^2 = 2 TiB)",
               {"chi": chi, "local_dims": dims, "flops_per_step": flops, "l2": f"local tensor {nbytes / 2**30:.2f} GiB > L2"},
               {"value": flops / e2e_s * 1e-12, "unit": "TFLOP/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes, "ms_per_step": e2e_s * 1e3},
               c["kernel_launches"], clk.summary(), roof, None, extra))


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5], help="BASELINE.json config (2 = headline)")
    ap.add_argument("--chi", type=int, default=4096)
    ap.add_argument("--nsites", type=int, default=100)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fused", action="store_true", help="multi-GPU: fused GEMM + peer-store reduce-scatter epilogue instead of the "
                    "NCCL collective")
    ap.add_argument("--no-region-step", action="store_true", help="skip the full region steps (extract + 3-matvec Lanczos + truncating insert)")
    ap.add_argument("--no-full-sweep", action="store_true", help="single GPU: skip the measured full 2-site DMRG sweep (about 2.5 minutes at chi=4096)")
    ap.add_argument("--full-sweep", action="store_true", help="multi-GPU: also run the measured full sweep (every rank in lock step)")
    ap.add_argument("--sweep-watchdog-s", type=float, default=900.0, help="print the line without the full sweep if the sweep has not "
                    "returned after this many seconds (measured: 138 s at chi=4096, N=100)")
    ap.add_argument("--region-steps", type=int, default=3, help="consecutive region steps timed in each sweep direction")
    ap.add_argument("--cutoff", type=float, default=0.0, help="inserter cutoff of the region steps (0: maxdim-limited; 1e-9: the reference's "
                    "timed_dmrg setting)")
    ap.add_argument("--opt", action="append", default=[], help="context option key=value (nsb_ctx_set_option), repeatable")
    ap.add_argument("--no-canonical", action="store_true", help="leave the synthetic state as filled (random tensors, gauge flag only) instead of "
                    "orthonormalising it to the benchmark bond with device QRs during set-up")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.config == 2:
        run_config2(args, rank, world, local_rank)
        return
    if rank != 0:
        return            # the other configs are single-GPU measurements
    {1: run_config1, 3: run_config3, 4: run_config4, 5: run_config5}[args.config](args)


def run_config3(args):
    from bench_qn import run_config3 as impl      # tools/bench_qn.py (Hubbard chain with QN conservation)
    impl(args, emit, _line, ClockSampler, roofline_from_profile, pinned_array)


if __name__ == "__main__":
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    main()
