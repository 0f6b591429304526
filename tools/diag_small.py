import sys, numpy as np
sys.path.insert(0, ".")
import networksolvers_b200 as ns
from bench import build_problem
chi, N = int(sys.argv[1]), int(sys.argv[2])
ctx = ns.default_context()
net, region = build_problem(chi, N, ctx)
net.extract(region)
ref = net.matvec_device(1, download=True)
for G in (2, 4):
    out, mode = net.shard_emulate(G)
    print("EMUL", G, mode, np.abs(out - ref).max() / np.abs(ref).max(), flush=True)
for reg in ([N // 2, N // 2 + 1], [N // 2 + 1, N // 2 + 2], [N // 2 + 2, N // 2 + 1]):
    net.extract(reg); val, _ = net.update_eigsolve(); ins = net.insert((0.0, 1, chi))
    print("STEP", reg, val, ins.newdim, flush=True)
