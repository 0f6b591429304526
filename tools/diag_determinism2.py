"""Repeated builds of the benchmark problem in one process: hashes of all site tensors and of the H_eff application."""
import sys, hashlib, numpy as np
sys.path.insert(0, ".")
import networksolvers_b200 as ns
from bench import build_problem
chi, N, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ctx = ns.default_context()
for opt in sys.argv[4:]:
    k, v = opt.split("="); ctx.set_option(k, int(v))
first = None
for rep in range(reps):
    net, region = build_problem(chi, N, ctx)
    net.extract(region)
    y = net.matvec_device(1, download=True)
    hs = net.to_host()
    hv = {v: hashlib.md5(np.ascontiguousarray(hs.tensors[v]).tobytes()).hexdigest()[:8] for v in hs.graph.vertices}
    hy = hashlib.md5(y.tobytes()).hexdigest()[:8]
    if first is None:
        first = (hv, hy, hs, y)
    bad = [v for v in hv if hv[v] != first[0][v]]
    msg = ""
    if bad:
        v = bad[0]
        d = np.abs(hs.tensors[v] - first[2].tensors[v])
        msg = f" first differing site {v} max abs diff {d.max():.3e} nnz {int((d > 0).sum())} of {d.size}"
    print(f"DET2 rep {rep} y {hy} {'same' if hy == first[1] else 'DIFF %.3e' % (np.abs(y - first[3]).max() / np.abs(y).max())} differing sites {bad[:8]}{msg}", flush=True)
    net.close()
