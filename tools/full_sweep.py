"""One measured 2-site DMRG sweep over all regions of the N-site Heisenberg chain at bond dimension chi through the public
driver, from a synthetic state orthonormalised on the device (random tensors + QR gauge walk), with per-phase device timers.
    python tools/full_sweep.py [chi] [nsites] [cutoff]"""
import json, sys, time
sys.path.insert(0, ".")
import networksolvers_b200 as ns

chi = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100
cutoff = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
ctx = ns.default_context()
g = ns.path_graph(N)
sites = ns.siteinds("S=1/2", g)
H = ns.ttno(ns.heisenberg(g), sites)
plan = ns.euler_sweep(g, nsites=2)
t0 = time.perf_counter()
net = ns.DeviceNetwork.synthetic(H, sites, chi, seed=1234, ctx=ctx, canonical=True)
net.extract(list(plan[0][0]))          # gauge walk to the first region + all environments (set-up, not timed)
ctx.synchronize()
setup_s = time.perf_counter() - t0
prob = ns.EigsolveProblem(net=net)
ctx.enable_timers(True); ctx.reset_timers(); ctx.reset_counters(); ctx.synchronize()
t0 = time.perf_counter()
E, _ = ns.dmrg(prob, nsweeps=1, nsites=2, inserter_kwargs=dict(trunc=dict(cutoff=cutoff, maxdim=chi)))
ctx.synchronize()
dt = time.perf_counter() - t0
c = ctx.counters()
print(json.dumps(dict(workload=f"2-site DMRG sweep, S=1/2 Heisenberg chain N={N}, chi={chi}, cutoff={cutoff}", full_sweep_s=dt,
                      regions=len(plan), setup_s=setup_s, energy=E, phase_ms=ctx.timers(), launches=int(c["kernel_launches"]),
                      matvecs=int(c["matvecs"]), gemm_tflop_executed=c["gemm_flops"] / 1e12, maxlinkdim=int(net.maxlinkdim()))), flush=True)
