"""Region-step timings for BASELINE configs 4 (2-site TDVP, complex, chi=1024) and 5 (tree, 1-site + expansion)."""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np
import networksolvers_b200 as ns

ctx = ns.default_context()
which = sys.argv[1] if len(sys.argv) > 1 else "4"
if which == "4":
    chi, N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024, 64
    for model in ("heisenberg", "ising"):
        g = ns.path_graph(N); sites = ns.siteinds("S=1/2", g)
        H = ns.ttno(ns.heisenberg(g) if model == "heisenberg" else ns.transverse_ising(g, 1.0, 1.0), sites)
        mid = N // 2
        net = ns.DeviceNetwork.synthetic(H, sites, chi, seed=7, dtype=np.complex128, ctx=ctx, ortho_region=[mid, mid + 1])
        t0 = time.perf_counter(); net.extract([mid, mid + 1]); ctx.synchronize(); t_ext = time.perf_counter() - t0
        flops = net.matvec_flops()
        net.matvec_device(2)
        ctx.tic(); net.matvec_device(5); ms = ctx.toc() / 5
        ctx.enable_timers(True); ctx.reset_timers()
        t0 = time.perf_counter()
        info = net.update_exp(-0.05j, solver="rk", order=4, nsites=2)
        ins = net.insert((1e-14, 1, chi), normalize=True)
        ctx.synchronize(); dt = time.perf_counter() - t0
        print(json.dumps(dict(config=4, model=model, chi=chi, N=N, dtype="c128", matvec_ms=ms, matvec_real_tflops=flops / ms * 1e-9,
                              setup_extract_s=t_ext, region_step_s=dt, nmatvec=info.nmatvec, newdim=ins.newdim, phases_ms=ctx.timers(),
                              jacobi_sweeps=ins.jacobi_sweeps)), flush=True)
        ctx.enable_timers(False)
        net.close()
else:
    chi = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    teeth = [6] * 10
    g = ns.named_comb_tree(teeth); sites = ns.siteinds("S=1/2", g)
    H = ns.ttno(ns.heisenberg(g), sites)
    v = (5, 1)                      # backbone vertex of degree 3
    net = ns.DeviceNetwork.synthetic(H, sites, chi, seed=7, dtype=np.float64, ctx=ctx, ortho_region=[v])
    t0 = time.perf_counter(); net.extract([v]); ctx.synchronize(); t_ext = time.perf_counter() - t0
    legs, dims = net.local_info()
    flops = net.matvec_flops()
    net.matvec_device(1)
    ctx.reset_counters()
    ctx.tic(); net.matvec_device(3); ms = ctx.toc() / 3
    c = ctx.counters()
    t0 = time.perf_counter()
    val, sinfo = net.update_eigsolve()
    ins = net.insert((1e-9, 1, chi))
    ctx.synchronize(); dt = time.perf_counter() - t0
    print(json.dumps(dict(config=5, graph="named_comb_tree 10x6 (60 sites)", chi=chi, local_dims=dims, matvec_ms=ms,
                          matvec_tflops=flops / ms * 1e-9, matvec_flops=flops, permute_bytes_per_matvec=c["permute_bytes"] / 3,
                          setup_extract_s=t_ext, region_step_s=dt, mem=ctx.mem_info())), flush=True)
