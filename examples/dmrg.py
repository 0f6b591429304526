"""examples/dmrg.jl of the reference through networksolvers_b200 (needs a B200).

    python examples/dmrg.py dmrg [--N 10] [--nsites 2] [--site-type "S=1"] [--conserve-qns]
    python examples/dmrg.py tree_dmrg
    python examples/dmrg.py sweep_loop_version
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import networksolvers_b200 as ns  # noqa: E402


def heisenberg_opsum(g):
    os_ = ns.OpSum()
    for u, v in g.edges:
        os_.add(1.0, "Sz", u, "Sz", v)
        os_.add(0.5, "S+", u, "S-", v)
        os_.add(0.5, "S-", u, "S+", v)
    return os_


def dmrg(N=10, nsites=2, site_type="S=1", conserve_qns=False, dry_run=False):
    """examples/dmrg.jl:10-47: chain of N sites, product start (even -> Up, odd -> Dn), 5 sweeps, cutoff 1e-12,
    maxdim [10, 40, 80, 160], density-matrix subspace expansion with factor 1.1."""
    g = ns.path_graph(N)
    s = ns.siteinds(site_type, g, conserve_qns=conserve_qns)
    H = ns.mpo(heisenberg_opsum(g), s)
    state = {v: ("Up" if j % 2 == 0 else "Dn") for j, v in enumerate(g.vertices, start=1)}
    psi = ns.product_state(s, state)
    trunc = dict(cutoff=1e-12, maxdim=[10, 40, 80, 160])
    extracter_kwargs = dict(trunc=trunc, subspace_algorithm="densitymatrix", expansion_factor=1.1)
    inserter_kwargs = dict(trunc=trunc)
    if dry_run:
        print(f"dmrg: N={N} site_type={site_type} operator link dimension {H.maxlinkdim()}")
        return None
    t0 = time.perf_counter()
    energy, gs_psi = ns.dmrg(H, psi, nsweeps=5, nsites=nsites, extracter_kwargs=extracter_kwargs,
                             inserter_kwargs=inserter_kwargs, outputlevel=1)
    print(f"  {time.perf_counter() - t0:.6f} seconds")
    print("Final energy = ", energy)
    if site_type == "S=1" and N == 10:
        print("Exact energy = -12.8945601")
    return energy


def tree_dmrg(dry_run=False):
    """examples/dmrg.jl:49-75: S = 1 Heisenberg model on a comb tree with three teeth of five sites."""
    c = ns.named_comb_tree([5, 5, 5])
    s = ns.siteinds("S=1", c)
    H = ns.ttno(heisenberg_opsum(c), s)
    psi = ns.random_state(s, 4, seed=1234)
    trunc = dict(cutoff=1e-9, maxdim=10)
    extracter_kwargs = dict(trunc=trunc, subspace_algorithm="densitymatrix", expansion_factor=1.2)
    if dry_run:
        print(f"tree_dmrg: {len(c.vertices)} vertices, operator link dimension {H.maxlinkdim()}")
        return None
    energy, gs_psi = ns.dmrg(H, psi, nsweeps=14, nsites=2, extracter_kwargs=extracter_kwargs,
                             inserter_kwargs=dict(trunc=trunc), outputlevel=2)
    print("Final energy = ", energy)
    return energy


def sweep_loop_version(dry_run=False):
    """examples/dmrg.jl:77-103: the sweep and region iterators driven by hand instead of through `dmrg`."""
    N = 10
    g = ns.path_graph(N)
    s = ns.siteinds("S=1", g)
    H = ns.mpo(heisenberg_opsum(g), s)
    psi = ns.random_state(s, 4, seed=1)
    nsweeps = 2
    trunc = dict(cutoff=1e-6, maxdim=[10, 20, 40, 100, 200])
    if dry_run:
        print("sweep_loop_version: host objects built")
        return None
    problem = ns.EigsolveProblem(state=psi, operator=H)
    sweeps = ns.sweep_iterator(problem, nsweeps, nsites=2, outputlevel=0, extracter_kwargs=dict(trunc=trunc), updater_kwargs={},
                               inserter_kwargs=dict(trunc=trunc))
    for sweep, region_iter in enumerate(sweeps, start=1):
        print(f"\nSweep {sweep}:")
        for region, _ in ns.region_tuples(region_iter):
            print(f"  Region {region}: energy = {ns.eigenvalue(ns.problem(region_iter)):.12f}")
        print(f"Done with sweep {sweep}")
    return ns.eigenvalue(ns.problem(region_iter))


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("which", nargs="?", default="dmrg", choices=["dmrg", "tree_dmrg", "sweep_loop_version"])
    ap.add_argument("--N", type=int, default=10)
    ap.add_argument("--nsites", type=int, default=2)
    ap.add_argument("--site-type", default="S=1")
    ap.add_argument("--conserve-qns", action="store_true", help="Sz conservation: block-sparse tensors on the device")
    ap.add_argument("--dry-run", action="store_true")
    a = ap.parse_args()
    if a.which == "dmrg":
        dmrg(a.N, a.nsites, a.site_type, a.conserve_qns, a.dry_run)
    elif a.which == "tree_dmrg":
        tree_dmrg(a.dry_run)
    else:
        sweep_loop_version(a.dry_run)
