import sys, time, json
sys.path.insert(0, ".")
import numpy as np
import networksolvers_b200 as ns
ctx = ns.default_context()
chi = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = ns.named_comb_tree([6] * 10); sites = ns.siteinds("S=1/2", g)
H = ns.ttno(ns.heisenberg(g), sites)
v = (5, 1)
net = ns.DeviceNetwork.synthetic(H, sites, chi, seed=7, dtype=np.float64, ctx=ctx, ortho_region=[v])
net.extract([v]); ctx.synchronize()
print(net.local_info(), "flops", net.matvec_flops())
for rep in range(3):
    ctx.reset_counters(); ctx.tic(); net.matvec_device(1); ms = ctx.toc(); c = ctx.counters()
    print("matvec ms", ms, {k: c[k] for k in ("kernel_launches", "gemm_calls", "gemm_flops", "permute_bytes")})
ctx.enable_timers(True); ctx.reset_timers()
t0 = time.perf_counter(); val, sinfo = net.update_eigsolve(); ctx.synchronize(); print("eigsolve s", time.perf_counter() - t0, val, sinfo.nmatvec, sinfo.krylovdim, ctx.timers())
t0 = time.perf_counter(); ins = net.insert((1e-9, 1, chi)); ctx.synchronize(); print("insert s", time.perf_counter() - t0, ins.newdim)
print(ctx.mem_info())
