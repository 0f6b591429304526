import numpy as np, sys
sys.path.insert(0, ".")
import networksolvers_b200 as ns
g = ns.path_graph(10); sites = ns.siteinds("S=1/2", g)
H = ns.ttno(ns.heisenberg(g), sites)
for cplx in (False, True):
    psi = ns.random_state(sites, 16, seed=5, dtype=complex if cplx else float)
    prob = ns.EigsolveProblem(state=psi, operator=H); net = prob.net; ctx = net.ctx
    def stage(name):
        c = ctx.counters(); print(f"  {name:28s} permute_bytes={c['permute_bytes']:8d} gemm={c['gemm_calls']:4d} launches={c['kernel_launches']}"); ctx.reset_counters()
    ctx.reset_counters()
    for region in ([4, 5], [5, 6], [6, 7], [7, 6], [6, 5], [5, 4]):
        print("region", region, "cplx", cplx)
        net.extract(region); stage("extract")
        net.matvec_device(1); stage("matvec")
        val, info = net.update_eigsolve(); stage("eigsolve")
        ins = net.insert((0.0, 1, 16)); stage("insert")
        print("   E=", val, "newdim", ins.newdim, "terr", ins.truncerr, "sweeps", ins.jacobi_sweeps)
