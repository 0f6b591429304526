"""Tournament Jacobi: slots per CTA (cluster width) sweep; wall time per nsb_factorize_host call."""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np
import networksolvers_b200 as ns
ctx = ns.default_context()
rng = np.random.default_rng(0)
for n in (48, 64, 100, 128, 166, 256):
    M = rng.standard_normal((n, n)) * np.exp(-0.1 * np.arange(n))[None, :]
    rec = dict(bench="dsmem_spc", n=n)
    for spc in (16, 8, 4, 2):
        ctx.set_option("jacobi_dsmem_spc", spc)
        ctx.factorize(M, cutoff=1e-12)
        ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            U, C, spec, info = ctx.factorize(M, cutoff=1e-12)
        ctx.synchronize()
        rec[f"spc{spc}_ms"] = round((time.perf_counter() - t0) / 5 * 1e3, 3)
        rec["sweeps"] = info["sweeps"]
    ctx.set_option("jacobi_dsmem_spc", 4)
    print(json.dumps(rec), flush=True)
