import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import networksolvers_b200 as ns
from helpers import neel
N = 10
g = ns.NamedGraph()
for j in range(1, N + 1): g.add_vertex(j)
for j in range(1, N): g.add_edge(j, j + 1)
g.add_vertex(0); g.add_edge(0, N // 2)
sites = ns.siteinds("S=1/2", g)
os_ = ns.OpSum()
for j in range(1, N):
    os_.add(1.0, "Sz", j, "Sz", j + 1); os_.add(0.5, "S+", j, "S-", j + 1); os_.add(0.5, "S-", j, "S+", j + 1)
H = ns.ttno(os_, sites)
psi0 = ns.product_state(sites, neel(g))
trunc = dict(cutoff=1e-10, maxdim=100)
def rcb(problem, region=None, **k):
    print("  region", region, getattr(problem, "eigenvalue", None), getattr(problem, "current_time", None), problem.last_info, flush=True)
print("DMRG", flush=True)
E, gs = ns.dmrg(H, psi0, nsweeps=2, nsites=2, inserter_kwargs=dict(trunc=trunc), region_callback=rcb)
gs_host = gs.to_host()
print("TDVP1", flush=True)
psi1 = ns.tdvp(H, gs_host, [0.0, 0.02], nsites=1, inserter_kwargs=dict(trunc=trunc), region_callback=rcb)
print("done", flush=True)
