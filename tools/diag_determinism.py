"""Is the set-up (Philox fill, gauge walk by blocked QR, environment builds) bitwise reproducible?  Two networks with the same
seed in one process; compares every site tensor and the H_eff application."""
import sys, numpy as np
sys.path.insert(0, ".")
import networksolvers_b200 as ns
from bench import build_problem
chi, N = int(sys.argv[1]), int(sys.argv[2])
ctx = ns.default_context()
for opt in sys.argv[3:]:
    k, v = opt.split("="); ctx.set_option(k, int(v))
res = []
for rep in range(2):
    net, region = build_problem(chi, N, ctx)
    net.extract(region)
    y = net.matvec_device(1, download=True)
    host = net.to_host()
    res.append((y, host))
    net.close()
ya, yb = res[0][0], res[1][0]
print("DET matvec identical:", np.array_equal(ya, yb), "max rel diff", np.abs(ya - yb).max() / np.abs(ya).max(), flush=True)
bad = []
for v in res[0][1].graph.vertices:
    a, b = res[0][1].tensors[v], res[1][1].tensors[v]
    if not np.array_equal(a, b):
        bad.append((v, float(np.abs(a - b).max())))
print("DET differing site tensors:", len(bad), bad[:12], flush=True)
