import sys, time
sys.path.insert(0, ".")
import numpy as np
import networksolvers_b200 as ns
ctx = ns.default_context()
rng = np.random.default_rng(0)
n = int(sys.argv[1])
U, _ = np.linalg.qr(rng.standard_normal((n, n))); V, _ = np.linalg.qr(rng.standard_normal((n, n)))
M = (U * np.exp(-12.0 * np.arange(n) / n)) @ V.T
for cap in [int(x) for x in sys.argv[2:]]:
    ctx.set_option("jacobi_inner_cap", cap)
    t0 = time.perf_counter()
    Uo, C, spec, info = ctx.factorize(M, cutoff=0.0, maxdim=n // 2)
    dt = time.perf_counter() - t0
    err = np.abs(Uo[:, :32].T @ Uo[:, :32] - np.eye(32)).max()
    print("cap", cap, "n", n, "time", round(dt, 3), "sweeps", info["sweeps"], "ortho", err, "spec err", np.abs(spec[:n // 2] - np.exp(-24.0 * np.arange(n // 2) / n)).max(), flush=True)
