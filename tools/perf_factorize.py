import json, sys, time
sys.path.insert(0, ".")
import numpy as np
import networksolvers_b200 as ns
ctx = ns.default_context()
rng = np.random.default_rng(0)
for n in (512, 1024, 2048, 4096, 8192):
    # spectrum like a DMRG two-site tensor: exponentially decaying singular values
    M = rng.standard_normal((n, n))
    t0 = time.perf_counter()
    ctx.reset_counters()
    U, C, spec, info = ctx.factorize(M, cutoff=0.0, maxdim=n // 2)
    dt = time.perf_counter() - t0
    c = ctx.counters()
    ortho = np.abs(U[:, :64].T @ U[:, :64] - np.eye(64)).max()
    print(json.dumps(dict(bench="factorize_blocked", n=n, s=dt, sweeps=info["sweeps"], gemm_tflop=c["gemm_flops"] / 1e12,
                          launches=c["kernel_launches"], ortho_err=ortho)), flush=True)
