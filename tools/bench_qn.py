"""BASELINE config 3 for `bench.py --config 3`: Hubbard chain N = 64 with (Nf, Sz) conservation, block-sparse sectors, bond
dimension 2048, 2-site DMRG region steps with the "densitymatrix" subspace expansion switched on (expansion_factor 1.1 as
examples/dmrg.jl:31).  SURVEY 8(d): synthetic block-sparse state with sector dimensions from a discretised Gaussian over
(Nf, Sz) summing to chi, seed 1234; dense-equivalent AND executed block flops are reported.

A step = one H_eff application on an interior bond through the sector-batched engine (csrc/bsparse.cu: environments, local
tensor and Krylov vectors as symmetry blocks, one grouped DMMA GEMM launch per contraction)."""
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def hubbard_link_sectors(N, j, chi, sigma_n=1.3, sigma_s=1.8):
    """Charges (Nf, 2 Sz) and dimensions of the sectors of the bond between sites j and j + 1 (j sites on the left) of a
    half-filled Sz = 0 Hubbard chain: Gaussian weights around (Nf, 2Sz) = (j, 0), capped by the number of basis states of
    the sector on either side, rounded to sum to min(chi, what both sides allow)."""
    def nstates(nsites, nf, sz2):          # basis states of `nsites` sites with Nf = nf, 2 Sz = sz2
        if (nf + sz2) % 2 or nf < 0:
            return 0
        nup, ndn = (nf + sz2) // 2, (nf - sz2) // 2
        if nup < 0 or ndn < 0 or nup > nsites or ndn > nsites:
            return 0
        return math.comb(nsites, nup) * math.comb(nsites, ndn)

    cand = []
    for nf in range(0, 2 * j + 1):
        for sz2 in range(-nf, nf + 1):
            cap = min(nstates(j, nf, sz2), nstates(N - j, N - nf, -sz2))
            if cap > 0:
                w = math.exp(-((nf - j) ** 2) / (2 * sigma_n ** 2) - (sz2 ** 2) / (2 * sigma_s ** 2))
                cand.append([nf, sz2, cap, w])
    target = min(chi, sum(c[2] for c in cand))
    dims = [0] * len(cand)
    remaining = target
    # water-filling: distribute by weight, respecting the caps
    active = list(range(len(cand)))
    for _ in range(64):
        if remaining <= 0 or not active:
            break
        wsum = sum(cand[i][3] for i in active)
        nxt = []
        given = 0
        for i in active:
            share = int(remaining * cand[i][3] / wsum)
            add = min(share, cand[i][2] - dims[i])
            dims[i] += add
            given += add
            if dims[i] < cand[i][2]:
                nxt.append(i)
        remaining -= given
        if given == 0:
            for i in sorted(nxt, key=lambda k: -cand[k][3]):
                if remaining <= 0:
                    break
                dims[i] += 1
                remaining -= 1
        active = [i for i in nxt if dims[i] < cand[i][2]]
    rows = []
    for c, d in zip(cand, dims):
        rows += [[c[0], c[1]]] * d
    return np.array(rows, dtype=np.int32).reshape(-1, 2)


def build_hubbard(ns, N, chi, ctx, seed=1234):
    g = ns.path_graph(N)
    sites = ns.siteinds("Electron", g, conserve_qns=True)
    H = ns.ttno(ns.hubbard(g, 1.0, 4.0), sites)
    from networksolvers_b200.models import SITE_CHARGES
    sc = np.array(SITE_CHARGES["Electron"], dtype=np.int32)
    verts = g.vertices
    link = {}
    for j in range(1, N):
        link[(verts[j - 1], verts[j])] = hubbard_link_sectors(N, j, chi)       # charges of the left part (u's side)
    net = ns.DeviceNetwork.synthetic_qn(H, sites, link, total=[N, 0], site_charges=sc, seed=seed, ctx=ctx)
    return net, g, {k: len(v) for k, v in link.items()}


def run_config3(args, emit, _line, ClockSampler, roofline_from_profile, pinned_array):
    import networksolvers_b200 as ns
    from networksolvers_b200 import _lib as L
    ctx = ns.default_context()
    N = args.nsites if args.nsites != 100 else 64
    chi = args.chi if args.chi != 4096 else 2048
    t0 = time.perf_counter()
    net, g, ldims = build_hubbard(ns, N, chi, ctx)
    mid = N // 2
    region = [mid, mid + 1]
    net.extract(region)
    ctx.synchronize()
    setup = time.perf_counter() - t0
    legs, dims = net.local_info()
    for _ in range(args.warmup):
        net.matvec_device(1)
    flops_exec, flops_dense = net.matvec_flops_executed(), net.matvec_flops()
    ctx.reset_counters()
    ctx.gemm_profile(True)
    marked = os.environ.get("NSB_PROFILE_TIMED") == "1"
    if marked:
        ctx.profiler(True)
    with ClockSampler(0) as clk:
        ctx.tic()
        net.matvec_device(args.steps)
        ms_total = ctx.toc()
    if marked:
        ctx.profiler(False)
    recs = ctx.gemm_profile_read()
    ctx.gemm_profile(False)
    c = ctx.counters()
    ms = ms_total / args.steps
    roof = roofline_from_profile(ctx, recs, ms_total)
    if roof:
        roof["note"] = ("grouped sector GEMMs (gemm_grouped_kernel): tiles of 128 x 128 over blocks of a few hundred rows, so part of "
                        "every tile is padding; achieved counts the block flops actually needed")
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
    # region steps: extract (+ densitymatrix expansion, expansion_factor 1.1) -> Lanczos -> truncating insert
    ex = dict(algorithm=L.NSB_EXPAND_DENSITYMATRIX, north_pass=1, expansion_factor=1.1, max_expand=L.INT64_MAX)
    ctx.enable_timers(True)
    steps, phases, newdims = [], [], []
    for r in range(max(args.region_steps, 1)):
        reg = [region[0] + r, region[1] + r]
        ctx.reset_timers()
        ctx.synchronize()
        t0 = time.perf_counter()
        net.extract(reg, (1e-12, 1, chi), ex)
        val, info = net.update_eigsolve()
        ins = net.insert((1e-12, 1, chi))
        ctx.synchronize()
        steps.append(time.perf_counter() - t0)
        phases.append(ctx.timers())
        newdims.append(int(ins.newdim))
    ctx.enable_timers(False)
    cpu = None
    if not args.no_cpu_baseline:
        # CPU arm: dense-storage matvec of the same shape through the oracle (the oracle has no block-sparse contraction; the
        # reference's ITensors QN path executes block flops) on a right-bond slab, all threads
        import bench as B
        s = B.CpuSlabMatvec(min(chi, 1024), nslabs=8, w=6, d=4)
        s.step()
        ts = [s.step() for _ in range(3)]
        cdt = float(np.mean(ts))
        cpu = {"value": s.flops / cdt * 1e-12, "unit": "TFLOP/s (dense-equivalent)", "cores": B.blas_threads(), "kind": "port",
               "sample": f"3 x one right-bond slab (1/8) of the dense-storage chi={min(chi, 1024)} d=4 w=6 matvec through the oracle's "
                         f"optimal_map (NumPy + {B.blas_name()}); the oracle contracts dense tensors, so this is a dense-equivalent rate"}
    extra = {"region_step_s": float(np.median(steps[1:] if len(steps) > 1 else steps)), "region_steps_s": steps,
             "region_phase_ms": {k: float(np.median([p[k] for p in (phases[1:] if len(phases) > 1 else phases)])) for k in phases[0]},
             "region_newdim": newdims, "setup_s": setup, "executed_flops_per_step": flops_exec, "dense_equivalent_flops_per_step": flops_dense,
             "executed_fraction": flops_exec / flops_dense, "dense_equivalent_tflops": flops_dense / ms * 1e-9,
             "bond_dimensions_requested": {"min": min(ldims.values()), "max": max(ldims.values())}, "maxlinkdim": int(net.maxlinkdim()),
             "hbm_pool_used_gib": ctx.mem_info()["pool_used"] / 2**30}
    emit(_line("heff_matvec_fp64_tflops", flops_exec / ms * 1e-9, "TFLOP/s", args, ms, True, "f64",
               f"Hubbard chain N={N} (t=1, U=4) with (Nf, Sz) conservation, block-sparse sectors, 2-site H_eff matvec on bond "
               f"({region[0]},{region[1]}), chi={chi}, d=4, w=6 + densitymatrix expansion in the region steps (BASELINE config 3)",
               {"chi": chi, "local_dims": dims, "flops_per_step": flops_exec,
                "flops_note": "value / e2e use the sector (block) flops actually executed; dense_equivalent_* is the dense-storage count",
                "l2": "block storage of L, R, theta and T1 exceeds L2 at chi = 2048"},
               {"value": flops_exec / e2e_s * 1e-12, "unit": "TFLOP/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
                "ms_per_step": e2e_s * 1e3},
               c["kernel_launches"], clk.summary(), roof, cpu, extra))
