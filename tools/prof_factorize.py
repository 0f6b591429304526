import sys, time
sys.path.insert(0, ".")
import numpy as np
import networksolvers_b200 as ns
ctx = ns.default_context()
rng = np.random.default_rng(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
kind = sys.argv[2] if len(sys.argv) > 2 else "gauss"
if kind == "gauss":
    M = rng.standard_normal((n, n))
else:
    U, _ = np.linalg.qr(rng.standard_normal((n, n))); V, _ = np.linalg.qr(rng.standard_normal((n, n)))
    M = (U * np.exp(-12.0 * np.arange(n) / n)) @ V.T
t0 = time.perf_counter()
U, C, spec, info = ctx.factorize(M, cutoff=0.0, maxdim=n // 2)
print(kind, n, time.perf_counter() - t0, info)
