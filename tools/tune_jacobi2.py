import sys, time
sys.path.insert(0, ".")
import numpy as np
import networksolvers_b200 as ns
ctx = ns.default_context()
rng = np.random.default_rng(0)
n = int(sys.argv[1])
U, _ = np.linalg.qr(rng.standard_normal((n, n))); V, _ = np.linalg.qr(rng.standard_normal((n, n)))
M = (U * np.exp(-12.0 * np.arange(n) / n)) @ V.T
for name, opts in (("nopivot", dict(jacobi_pivot=0)), ("pivot", dict(jacobi_pivot=1)), ("noprecond", dict(jacobi_precondition=0))):
    for k, v in dict(jacobi_pivot=0, jacobi_precondition=1).items(): ctx.set_option(k, v)
    for k, v in opts.items(): ctx.set_option(k, v)
    ctx.enable_timers(False)
    t0 = time.perf_counter()
    Uo, C, spec, info = ctx.factorize(M, cutoff=0.0, maxdim=n // 2)
    dt = time.perf_counter() - t0
    print(name, "n", n, "time", round(dt, 3), "sweeps", info["sweeps"], "ortho", np.abs(Uo[:, :32].T @ Uo[:, :32] - np.eye(32)).max(), flush=True)
