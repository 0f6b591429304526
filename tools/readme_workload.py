"""The reference README's timing-note workload (README.md:119-187): N=100 S=1/2 Heisenberg chain, 5 sweeps,
cutoff 1e-9, maxdim [10,40,80,160]; reference: 2-site 5.8 s, 1-site+densitymatrix(maxdim 4.. ) 9.8 s (hardware unstated)."""
import sys, time, json
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import networksolvers_b200 as ns
from helpers import neel
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
TIMERS = len(sys.argv) > 2 and sys.argv[2] == 'timers'   # per-phase device timers (adds event syncs)
g = ns.path_graph(N); sites = ns.siteinds("S=1/2", g)
H = ns.ttno(ns.heisenberg(g), sites)
psi0 = ns.product_state(sites, neel(g))
trunc = dict(cutoff=1e-9, maxdim=[10, 40, 80, 160])
ctx = ns.default_context()
for a in sys.argv[3:]:   # context options, e.g. eigh_min_n=256
    k, v = a.split('='); ctx.set_option(k, int(v))
for nsites, ek in ((2, {}), (1, dict(trunc=trunc, subspace_algorithm="densitymatrix", expansion_factor=1.5))):
    ctx.reset_counters()
    ctx.enable_timers(TIMERS); ctx.reset_timers()
    t0 = time.perf_counter()
    E, psi = ns.dmrg(H, psi0, nsweeps=5, nsites=nsites, extracter_kwargs=ek, inserter_kwargs=dict(trunc=trunc), outputlevel=1)
    dt = time.perf_counter() - t0
    c = ctx.counters()
    print(json.dumps(dict(workload="README timing note", N=N, nsites=nsites, seconds=dt, energy=E, maxlinkdim=psi.maxlinkdim(),
                          launches=c["kernel_launches"], matvecs=c["matvecs"], counters=c, phase_ms=ctx.timers() if TIMERS else None, reference_s=5.8 if nsites == 2 else 9.8)), flush=True)
