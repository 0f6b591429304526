"""Generates tests/golden/oracle_anchors.json: per-sweep energies, truncation errors, bond dimensions, TDVP fidelities
and fitting overlaps of the reference algorithm, from the NumPy oracle (`oracle/`) together with independent exact
diagonalisation.  The Julia reference cannot run in this environment (no julia, un-vendored dependencies), so these
are *restatement* values pinned by ED and by the constants the reference ships (examples/dmrg.jl:42,
test/dmrg/test_tree_dmrg.jl:53); they freeze the oracle against drift and give the GPU tests fixed vectors.

    python tools/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import sweep as osw                                    # noqa: E402
from oracle.ed import ed_ground_state, ed_time_evolution, state_vector   # noqa: E402
from oracle.graph import build_tree, path_graph                     # noqa: E402
from oracle.models import heisenberg_opsum, product_ttn, spin_ops, ttno   # noqa: E402


def neel(g, st, even_up=True):
    out = {}
    for j, v in enumerate(g.vertices, start=1):
        up = (j % 2 == 0) if even_up else (j % 2 == 1)
        out[v] = st["Up"] if up else st["Dn"]
    return out


def dmrg_case(name, g, site_type, nsweeps, nsites, trunc, extracter_kwargs=None):
    d, ops, st = spin_ops(site_type)
    H = ttno(heisenberg_opsum(g), g, ops)
    psi0 = product_ttn(g, d, neel(g, st))
    rec = {"E": [], "maxdim": []}
    osw.COUNTERS.clear()

    def cb(region_iter, **k):
        rec["E"].append(float(region_iter.problem.eigenvalue))
        rec["maxdim"].append(int(region_iter.problem.state.maxlinkdim()))

    ek = dict(extracter_kwargs or {})
    E, psi = osw.dmrg(H, psi0, nsweeps=nsweeps, nsites=nsites, extracter_kwargs=ek, inserter_kwargs=dict(trunc=trunc), sweep_callback=cb)
    terr = [float(t) for t in osw.COUNTERS.get("truncerrs", [])]
    E0 = float(np.atleast_1d(ed_ground_state(heisenberg_opsum(g), g, ops)[0])[0]) if len(g.vertices) <= 14 else None
    return {"name": name, "site_type": site_type, "nsweeps": nsweeps, "nsites": nsites, "trunc": trunc,
            "extracter_kwargs": {k: v for k, v in ek.items() if k != "trunc"}, "energies": rec["E"], "maxlinkdims": rec["maxdim"],
            "truncerr_sum": float(np.sum(terr)), "truncerr_max": float(np.max(terr)) if terr else 0.0, "n_truncerr": len(terr),
            "ed_energy": E0}


def main():
    out = {"generator": "tools/make_golden.py (oracle restatement + exact diagonalisation; the Julia reference is not runnable here)",
           "cases": {}}
    g = path_graph(10)
    out["cases"]["dmrg_s1_n10_2site"] = dmrg_case("examples/dmrg.jl:26-43 shape", g, "S=1", 5, 2, dict(cutoff=1e-12, maxdim=[10, 40, 80, 160]))
    out["cases"]["dmrg_shalf_n12_2site_eigen"] = dmrg_case("cutoff 1e-9 (eigen route)", path_graph(12), "S=1/2", 3, 2, dict(cutoff=1e-9, maxdim=[10, 40]))
    out["cases"]["dmrg_shalf_n14_2site"] = dmrg_case("maxdim-limited", path_graph(14), "S=1/2", 3, 2, dict(cutoff=1e-12, maxdim=[10, 20, 40]))
    out["cases"]["dmrg_tree_2site"] = dmrg_case("test/dmrg/test_tree_dmrg.jl:15-53", build_tree(3, 3), "S=1/2", 5, 2, dict(cutoff=1e-5, maxdim=[10, 20, 40]))
    tr = dict(cutoff=1e-12, maxdim=[10, 40, 80, 160])
    out["cases"]["dmrg_s1_n10_1site_expansion"] = dmrg_case("examples/dmrg.jl:26-36 (1-site + densitymatrix expansion 1.1)", g, "S=1", 5, 1, tr,
                                                            dict(trunc=tr, subspace_algorithm="densitymatrix", expansion_factor=1.1))
    # 2-site TDVP, examples/quench_evolution.jl shape
    from oracle.local_solvers import runge_kutta_solver
    g8 = path_graph(8)
    d, ops, st = spin_ops("S=1/2")
    H = ttno(heisenberg_opsum(g8), g8, ops, dtype=complex)
    psi0 = product_ttn(g8, d, neel(g8, st, even_up=False), dtype=complex)
    tp = [0.0, 0.05, 0.1, 0.15, 0.2]
    ik = dict(trunc=dict(maxdim=5000, cutoff=1e-14), normalize=True)
    tdv = {}
    for order in (2, 4):
        po = osw.tdvp(H, psi0, tp, nsites=2, tdvp_order=order, updater_kwargs=dict(solver=runge_kutta_solver, order=4), inserter_kwargs=ik)
        v = state_vector(po)
        vx = ed_time_evolution(heisenberg_opsum(g8), g8, ops, state_vector(psi0), tp, normalize=True)
        sz1 = float(np.real(np.vdot(v, np.kron(ops["Sz"], np.eye(2 ** 7)) @ v)))
        tdv[str(order)] = {"one_minus_fidelity_vs_expm": float(1 - abs(np.vdot(vx, v))), "maxlinkdim": int(po.maxlinkdim()), "sz_site1": sz1}
    out["cases"]["tdvp_n8_2site_rk4"] = {"time_points": tp, "orders": tdv}
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    with open(os.path.join(ROOT, "tests", "golden", "oracle_anchors.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1)[:1500])


if __name__ == "__main__":
    main()
