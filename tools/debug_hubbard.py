import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import networksolvers_b200 as ns
from helpers import to_oracle_ttn, SweepRecorder
from oracle import sweep as osw
g = ns.path_graph(6); sites = ns.siteinds("Electron", g)
H = ns.ttno(ns.hubbard(g, 1.0, 4.0), sites)
psi0 = ns.product_state(sites, {v: ("Up" if v % 2 else "Dn") for v in g.vertices})
trunc = dict(cutoff=1e-10, maxdim=[10, 20, 60])
ek = dict(trunc=trunc, subspace_algorithm="densitymatrix", expansion_factor=1.1)
log = []
def rcb(problem, region=None, **k):
    log.append((tuple(region), problem.eigenvalue, problem.last_info.get("expanded"), problem.last_info.get("newdim"), problem.last_truncerr))
E, psi = ns.dmrg(H, psi0, nsweeps=2, nsites=2, extracter_kwargs=ek, inserter_kwargs=dict(trunc=trunc), region_callback=rcb)
olog = []
def orcb(problem, region=None, **k):
    olog.append((tuple(region), problem.eigenvalue, problem.state.linkdims()))
osw.COUNTERS.clear()
Eo, psio = osw.dmrg(to_oracle_ttn(H, True), to_oracle_ttn(psi0), nsweeps=2, nsites=2, extracter_kwargs=ek, inserter_kwargs=dict(trunc=trunc), region_callback=orcb)
for a, b in zip(log, olog):
    print(a, "| oracle", b[1], sorted(b[2].values()))
