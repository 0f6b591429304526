"""CPU checks of the example scripts (examples/*.py: counterparts of the reference's examples/*.jl) and of the host-side pieces
they rely on: every script builds its host objects (`--dry-run` stops before the first device call), `expect` agrees with the
dense state, and a problem object may bring its own `region_iterator_action` (the reference's dispatch on the problem type,
examples/timed_dmrg/timed_eigsolve.jl:41-75)."""
import os
import subprocess
import sys
from functools import reduce

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("args", [["dmrg.py", "dmrg"], ["dmrg.py", "tree_dmrg"], ["dmrg.py", "sweep_loop_version"],
                                  ["quench_evolution.py"], ["tdvp.py", "tdvp"], ["tdvp.py", "test_tdvp"], ["fitting.py"],
                                  ["timed_dmrg.py"]])
def test_example_builds_its_host_objects(args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", args[0])] + args[1:] + ["--dry-run"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip()


def test_expect_matches_dense_state():
    import networksolvers_b200 as ns
    for g in (ns.path_graph(6), ns.named_comb_tree([2, 3, 2])):
        s = ns.siteinds("S=1/2", g)
        V = g.vertices
        for dt in (float, complex):
            psi = ns.random_state(s, 4, seed=3, dtype=dt)
            vec = psi.to_dense()
            for v in (V[0], V[len(V) // 2], V[-1]):
                for name in ("Sz", "S+"):
                    O = np.asarray(s.type.op(name))
                    full = reduce(np.kron, [O if u == v else np.eye(2) for u in V])
                    ref = np.vdot(vec, full @ vec) / np.vdot(vec, vec)
                    assert abs(ns.expect(psi, name, v, s) - ref) < 1e-12


def test_inner_matches_dense_states():
    import networksolvers_b200 as ns
    for g in (ns.path_graph(6), ns.named_comb_tree([2, 3, 2])):
        s = ns.siteinds("S=1/2", g)
        for dt in (float, complex):
            a, b = ns.random_state(s, 4, seed=3, dtype=dt), ns.random_state(s, 3, seed=5, dtype=dt)
            ref = np.vdot(a.to_dense(), b.to_dense())
            assert abs(ns.inner(a, b) - ref) < 1e-12 * max(1.0, abs(ref))


def test_problem_types_may_bring_their_own_region_action():
    """RegionIterator calls `problem.region_iterator_action` when the problem object defines one, with the region's keyword
    pack, and keeps whatever it returns as the problem -- no device needed to see the dispatch."""
    import networksolvers_b200 as ns

    class Wrapped:
        def __init__(self):
            self.calls = []

        def region_iterator_action(self, region_iterator, *, sweep, nsites, **kws):
            self.calls.append((ns.current_region(region_iterator), sweep, nsites, sorted(kws)))
            return self

    plan = [([1, 2], dict(sweep=1, nsites=2, extracter_kwargs={}, outputlevel=0)), ([2, 3], dict(sweep=1, nsites=2, outputlevel=0))]
    prob = Wrapped()
    it = ns.RegionIterator(prob, plan)
    seen = [ns.current_region(r) for r in it]
    assert seen == [[1, 2], [2, 3]]
    assert prob.calls == [([1, 2], 1, 2, ["extracter_kwargs", "outputlevel"]), ([2, 3], 1, 2, ["outputlevel"])]
    assert it.problem is prob and ns.is_last_region(it)
