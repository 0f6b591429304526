"""examples/tdvp.jl of the reference through networksolvers_b200 (needs a B200).

    python examples/tdvp.py tdvp [--N 10] [--total-time 1.0] [--time-step 0.1]
    python examples/tdvp.py test_tdvp [--N 6] [--total-time 0.5] [--time-step 0.02] [--tdvp-order 2]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import networksolvers_b200 as ns  # noqa: E402
from quench_evolution import dense_hamiltonian  # noqa: E402


def inner(a, b):
    return ns.inner(a.to_host(), b.to_host())


def tdvp(N=10, total_time=1.0, time_step=0.1, dry_run=False):
    """examples/tdvp.jl:9-63: the same random S = 1 state evolved with 2-site RK4, 1-site RK4 and 1-site Krylov exponentiate."""
    g = ns.path_graph(N)
    s = ns.siteinds("S=1", g)
    H = ns.mpo(ns.heisenberg(g), s)
    psi0 = ns.random_state(s, 60, seed=1)
    tdvp_order, outputlevel = 2, 0
    inserter_kwargs = dict(trunc=dict(maxdim=60, cutoff=1e-12), normalize=True)
    time_range = list(np.arange(0.0, total_time + 0.5 * time_step, time_step))
    if dry_run:
        print(f"tdvp: N={N}, {len(time_range)} time points, maxlinkdim(psi0) = {psi0.maxlinkdim()}")
        return None
    updater_kwargs = dict(solver=ns.runge_kutta_solver, order=4)
    print("Calling TDVP with RK4 solver and nsites=2 (res_2site)")
    res_2site = ns.tdvp(H, psi0, time_range, nsites=2, inserter_kwargs=inserter_kwargs, updater_kwargs=updater_kwargs,
                        tdvp_order=tdvp_order, outputlevel=outputlevel)
    print("maxlinkdim(res_2site) =", res_2site.maxlinkdim())
    print("Calling TDVP with RK4 solver (res_rk4)")
    res_rk4 = ns.tdvp(H, psi0, time_range, inserter_kwargs=inserter_kwargs, updater_kwargs=updater_kwargs,
                      tdvp_order=tdvp_order, outputlevel=outputlevel)
    print("maxlinkdim(res_rk4) =", res_rk4.maxlinkdim())
    print("inner(res_rk4, res_2site) =", inner(res_rk4, res_2site))
    print("Calling TDVP with exponentiate solver (res_1site)")
    res_1site = ns.tdvp(H, psi0, time_range, nsites=2, inserter_kwargs=inserter_kwargs, tdvp_order=tdvp_order,
                        outputlevel=outputlevel)
    print("inner(res_1site, res_rk4) =", inner(res_1site, res_rk4))
    print("inner(res_1site, res_2site) =", inner(res_1site, res_2site))
    return res_2site


def test_tdvp(N=6, total_time=0.5, time_step=0.02, tdvp_order=2, dry_run=False):
    """examples/tdvp.jl:65-149: <Sz_j>(t) from 2-site TDVP (sweep callback) against exact diagonalisation."""
    g = ns.path_graph(N)
    V = g.vertices
    s = ns.siteinds("S=1", g)
    os_ = ns.heisenberg(g)
    H = ns.mpo(os_, s)
    psi0 = ns.random_state(s, 30, seed=1)
    time_range = list(np.arange(0.0, total_time + 0.5 * time_step, time_step))
    nsweeps = len(time_range) - 1
    szs_tdvp = np.zeros((nsweeps, N))

    def sweep_callback(problem, *, sweep, **kws):
        host = problem.state.to_host()
        szs_tdvp[sweep - 1, :] = [ns.expect(host, "Sz", v, s) for v in V]

    inserter_kwargs = dict(trunc=dict(maxdim=40, cutoff=1e-10), normalize=True)
    if dry_run:
        print(f"test_tdvp: N={N}, {nsweeps} sweeps")
        return None
    psi_tdvp = ns.tdvp(H, psi0, time_range, nsites=2, extracter_kwargs={}, inserter_kwargs=inserter_kwargs, outputlevel=0,
                       sweep_callback=sweep_callback, tdvp_order=tdvp_order)
    print("\nResult from TDVP:")
    print(szs_tdvp)
    print("norm(psi_tdvp) =", psi_tdvp.norm())
    print("maxlinkdim(psi_tdvp) =", psi_tdvp.maxlinkdim())
    # ED
    Hx = dense_hamiltonian(os_, s)
    w, U = np.linalg.eigh(Hx)
    step = (U * np.exp(-1j * w * time_step)) @ U.conj().T
    psix = psi0.to_dense().astype(complex)
    psix /= np.linalg.norm(psix)
    d = s.dim
    Sz = np.asarray(s.type.op("Sz"))
    szs_ed = np.zeros((nsweeps, N))
    for sweep in range(nsweeps):
        psix = step @ psix
        psix /= np.linalg.norm(psix)
        t = psix.reshape([d] * N)
        for j in range(N):
            szs_ed[sweep, j] = np.vdot(t, np.moveaxis(np.tensordot(Sz, t, axes=(1, j)), 0, j)).real
    print("\nResult from ED:")
    print(szs_ed)
    print("fidelity =", abs(np.vdot(psi_tdvp.to_host().to_dense(), psix)))
    err = np.abs(szs_ed - szs_tdvp)
    i, j = np.unravel_index(np.argmax(err), err.shape)
    print("\nnorm(szs_ed - szs_tdvp) =", np.linalg.norm(szs_ed - szs_tdvp))
    print("Largest error (%.3E) at i,j=%d,%d" % (err[i, j], i + 1, j + 1))
    print("   TDVP value = %.10f" % szs_tdvp[i, j])
    print("     ED value = %.10f" % szs_ed[i, j])
    return err.max()


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("which", nargs="?", default="tdvp", choices=["tdvp", "test_tdvp"])
    ap.add_argument("--N", type=int, default=None)
    ap.add_argument("--total-time", type=float, default=None)
    ap.add_argument("--time-step", type=float, default=None)
    ap.add_argument("--tdvp-order", type=int, default=2)
    ap.add_argument("--dry-run", action="store_true")
    a = ap.parse_args()
    kw = {k: v for k, v in dict(N=a.N, total_time=a.total_time, time_step=a.time_step).items() if v is not None}
    if a.which == "tdvp":
        tdvp(dry_run=a.dry_run, **kw)
    else:
        test_tdvp(tdvp_order=a.tdvp_order, dry_run=a.dry_run, **kw)
