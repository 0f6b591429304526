"""Times the experimental bulge-chasing kernel (csrc/sbr.cu) on a random symmetric band matrix.
    NSB_DEBUG_EIGH=1 python tools/perf_sbr_chase.py [n] [b]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import networksolvers_b200 as ns
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
b = int(sys.argv[2]) if len(sys.argv) > 2 else 64
ctx = ns.default_context()
rng = np.random.default_rng(0)
ab = np.zeros((2 * b + 1, n), order="F")
ab[:b + 1, :] = rng.standard_normal((b + 1, n))
for j in range(n - b, n):
    ab[n - j:, j] = 0.0          # entries below the last row
for rep in range(2):
    t0 = time.perf_counter()
    out, V2, tau2 = ctx.sbr_chase(ab, b)
    print("call", rep, "wall", round(time.perf_counter() - t0, 3), "s, below sub-diagonal max", float(np.abs(out[2:, :]).max()), flush=True)
