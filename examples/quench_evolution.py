"""examples/quench_evolution.jl of the reference through networksolvers_b200 (needs a B200): 2-site TDVP quench of the S = 1/2
Heisenberg chain from the Neel state, <Sz> at the centre recorded by a sweep callback, ED fidelity for N <= 8.

    python examples/quench_evolution.py [--N 8] [--total-time 4.0] [--time-step 0.05] [--tdvp-order 4]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import networksolvers_b200 as ns  # noqa: E402


def print_linkdims(psi):
    for e, d in psi.linkdims().items():
        print(f"  {e}: {d}")


def ed_time_evolution(Hdense, v0, time_points, normalize=True):
    """test/utilities/simple_ed_methods.jl: exp(-i H t) |psi0> by exact diagonalisation (small N only)."""
    w, U = np.linalg.eigh(Hdense)
    t = time_points[-1] - time_points[0]
    v = U @ (np.exp(-1j * w * t) * (U.conj().T @ v0))
    return v / np.linalg.norm(v) if normalize else v


def dense_hamiltonian(os_, sites):
    from functools import reduce
    V, d, op = sites.graph.vertices, sites.dim, sites.type.op
    H = np.zeros((d ** len(V),) * 2, dtype=complex)
    for term in os_.terms:
        mats = {v: np.asarray(op(nm)) for nm, v in zip(term[1::2], term[2::2])}
        H += term[0] * reduce(np.kron, [mats.get(v, np.eye(d)) for v in V])
    return H


def quench(N=8, total_time=4.0, time_step=0.05, cutoff=1e-14, maxdim=5000, tdvp_order=4, dry_run=False):
    g = ns.path_graph(N)
    sites = ns.siteinds("S=1/2", g)
    os_ = ns.heisenberg(g)
    H = ns.mpo(os_, sites)
    V = g.vertices
    state = {v: ("Up" if j % 2 == 1 else "Dn") for j, v in enumerate(V, start=1)}
    psi0 = ns.product_state(sites, state)
    time_range = list(np.arange(0.0, total_time + 0.5 * time_step, time_step))
    centre = V[N // 2 - 1]
    szs = [ns.expect(psi0, "Sz", centre, sites)]

    def sweep_callback(problem, **kws):        # `problem.state` as in the reference's callback
        szs.append(ns.expect(problem.state.to_host(), "Sz", centre, sites))

    extracter_kwargs = dict(subspace_algorithm="densitymatrix", expansion_factor=1.2, max_expand=4)
    updater_kwargs = dict(solver=ns.runge_kutta_solver, order=4)
    inserter_kwargs = dict(trunc=dict(maxdim=maxdim, cutoff=cutoff), normalize=True)
    if dry_run:
        print(f"quench: N={N}, {len(time_range)} time points, <Sz>_centre(0) = {szs[0]:+.3f}")
        return None
    print("Calling TDVP")
    psif = ns.tdvp(H, psi0, time_range, nsites=2, sweep_callback=sweep_callback, extracter_kwargs=extracter_kwargs,
                   updater_kwargs=updater_kwargs, inserter_kwargs=inserter_kwargs, outputlevel=0, tdvp_order=tdvp_order)
    print_linkdims(psif)
    fname = f"szs_tdvp_N{N}.dat"
    print(f'Writing file "{fname}"')
    np.savetxt(fname, np.column_stack([time_range[: len(szs)], szs]))
    if N <= 8:
        print("Using ED to check")
        psix = ed_time_evolution(dense_hamiltonian(os_, sites), psi0.to_dense().astype(complex), time_range)
        v = psif.to_host().to_dense()
        print("Fidelity <psi_exact|psi_tdvp> = %.12f" % abs(np.vdot(psix, v)))
    return szs


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--N", type=int, default=8)
    ap.add_argument("--total-time", type=float, default=4.0)
    ap.add_argument("--time-step", type=float, default=0.05)
    ap.add_argument("--tdvp-order", type=int, default=4)
    ap.add_argument("--dry-run", action="store_true")
    a = ap.parse_args()
    quench(a.N, a.total_time, a.time_step, tdvp_order=a.tdvp_order, dry_run=a.dry_run)
