"""examples/timed_dmrg/*.jl of the reference through networksolvers_b200 (needs a B200): a wrapper problem type that times the
three hooks of every region step -- the extension seam of the reference (dispatch of `region_iterator_action!` on the problem
type), which is also where the GPU path plugs in.

    python examples/timed_dmrg.py [--N 100] [--nsites 2] [--site-type "S=1"]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import networksolvers_b200 as ns  # noqa: E402


class TimedEigsolveProblem:
    """Wraps an EigsolveProblem; accumulates wall time per hook (device work included: each hook ends synchronised)."""

    def __init__(self, eigprob):
        self.eigprob = eigprob
        self.reset_timings()

    def reset_timings(self):
        self.extracter_time = self.updater_time = self.inserter_time = 0.0

    # what the drivers and printers ask of a problem
    eigenvalue = property(lambda self: self.eigprob.eigenvalue)
    state = property(lambda self: self.eigprob.state)
    operator = property(lambda self: self.eigprob.operator)
    net = property(lambda self: self.eigprob.net)

    def region_iterator_action(self, region_iterator, *, extracter_kwargs=None, updater_kwargs=None, inserter_kwargs=None,
                               sweep, **kws):
        sync = self.eigprob.net.ctx.synchronize
        t0 = time.perf_counter()
        prob, local_state = ns.extracter(self.eigprob, region_iterator, **{**(extracter_kwargs or {}), "sweep": sweep, **kws})
        sync()
        t1 = time.perf_counter()
        prob, local_state = ns.updater(prob, local_state, region_iterator, **{**(updater_kwargs or {}), **kws})
        sync()
        t2 = time.perf_counter()
        prob = ns.inserter(prob, local_state, region_iterator, **{"sweep": sweep, **(inserter_kwargs or {}), **kws})
        sync()
        t3 = time.perf_counter()
        self.extracter_time += t1 - t0
        self.updater_time += t2 - t1
        self.inserter_time += t3 - t2
        self.eigprob = prob
        return self


def timed_eigsolve_sweep_printer(region_iterator, *, outputlevel, **kws):
    problem = region_iterator.problem
    ns.eigsolve_sweep_printer(region_iterator, outputlevel=outputlevel, **kws)
    if outputlevel >= 1:
        print("  Extracter time = %.3f s" % problem.extracter_time)
        print("  Updater time = %.3f s" % problem.updater_time)
        print("  Inserter time = %.3f s" % problem.inserter_time)
        problem.reset_timings()
        print(flush=True)


def timed_eigsolve(H, psi0, *, sweep_printer=timed_eigsolve_sweep_printer, **kws):
    eigprob = ns.EigsolveProblem(state=psi0, operator=H)
    return ns.eigsolve(TimedEigsolveProblem(eigprob), sweep_printer=sweep_printer, **kws)


timed_dmrg = timed_eigsolve


def main(N=100, nsites=2, site_type="S=1", dry_run=False):
    g = ns.path_graph(N)
    s = ns.siteinds(site_type, g)
    H = ns.mpo(ns.heisenberg(g), s)
    state = {v: ("Up" if j % 2 == 0 else "Dn") for j, v in enumerate(g.vertices, start=1)}
    psi = ns.product_state(s, state)
    trunc = dict(cutoff=1e-9, maxdim=[10, 40, 80, 160])
    if dry_run:
        print(f"timed_dmrg: N={N} site_type={site_type}")
        return None
    t0 = time.perf_counter()
    energy, gs_psi = timed_dmrg(H, psi, nsweeps=4, nsites=nsites, extracter_kwargs={}, inserter_kwargs=dict(trunc=trunc), outputlevel=1)
    print(f"  {time.perf_counter() - t0:.6f} seconds")
    print("Final energy = ", energy)
    if site_type == "S=1" and N == 10:
        print("Exact energy = -12.8945601")
    return energy


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--N", type=int, default=100)
    ap.add_argument("--nsites", type=int, default=2)
    ap.add_argument("--site-type", default="S=1")
    ap.add_argument("--dry-run", action="store_true")
    a = ap.parse_args()
    main(a.N, a.nsites, a.site_type, a.dry_run)
