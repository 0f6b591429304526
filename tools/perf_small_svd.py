"""Small truncating factorisations (the latency-bound regime): wall time per call through nsb_factorize_host for the cluster
Jacobi (one launch) and the launch-per-round kernels it replaces."""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np
import networksolvers_b200 as ns
ctx = ns.default_context()
rng = np.random.default_rng(0)
for n in (16, 32, 48, 64, 80, 100, 128, 166, 256):
    M = rng.standard_normal((n, n)) * np.exp(-0.1 * np.arange(n))[None, :]
    rec = dict(bench="small_svd", n=n)
    for name, mx, dx in (("cluster", 256, 0), ("rounds", 0, 0), ("dsmem", 0, 256)):
        ctx.set_option("jacobi_cluster_max_n", mx)
        ctx.set_option("jacobi_dsmem_max_n", dx)
        ctx.factorize(M, cutoff=1e-12)
        ctx.synchronize()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            U, C, spec, info = ctx.factorize(M, cutoff=1e-12)
        ctx.synchronize()
        rec[name + "_ms"] = (time.perf_counter() - t0) / reps * 1e3
        rec[name + "_sweeps"] = info["sweeps"]
    ctx.set_option("jacobi_cluster_max_n", 112)
    ctx.set_option("jacobi_dsmem_max_n", 256)
    print(json.dumps(rec), flush=True)
