import sys, numpy as np
sys.path.insert(0, ".")
import networksolvers_b200 as ns
from bench import build_problem
chi, N = int(sys.argv[1]), int(sys.argv[2])
ctx = ns.default_context()
net, region = build_problem(chi, N, ctx)
net.extract(region)
ref = net.matvec_device(1, download=True)
for G in (2, 8):
    out, mode = net.shard_emulate(G)
    d = np.abs(out - ref)
    per = d.shape[-1] // G
    print("EMUL", chi, N, G, mode, d.max() / np.abs(ref).max(), [float(d[..., r*per:(r+1)*per].max()) for r in range(G)], flush=True)
