"""CPU tests of the host-side sweep driver (the mirror of src/iterators.jl, src/sweep_solve.jl, src/eigsolve.jl:45-76,
src/applyexp.jl:18-103, src/inserter.jl, src/extracter.jl) with a recording stand-in for the device network: which hook is
called when, with which per-sweep parameters -- the same sequence the oracle's drivers (and the reference's) produce.
No arithmetic happens here; the arithmetic behind each call is what the `-m gpu` parity tests check."""
import sys
from types import SimpleNamespace

import pytest

import networksolvers_b200 as ns
from networksolvers_b200 import solvers as S
from networksolvers_b200 import _lib as L


class RecordingNet:
    """Stands in for DeviceNetwork: records the three hook calls per region."""

    def __init__(self, graph):
        self.graph = graph
        self.log = []

    def extract(self, region, trunc=None, expand=None):
        self.log.append(("extract", tuple(region), trunc, None if expand is None else dict(expand)))
        return SimpleNamespace(expanded=0, env_builds=1, qr_steps=1)

    def update_eigsolve(self, **kw):
        self.log.append(("eigsolve", kw))
        return -1.0 - 0.001 * len(self.log), SimpleNamespace(nmatvec=kw["krylovdim"], residual=0.0)

    def update_exp(self, t, **kw):
        self.log.append(("exp", complex(t), kw))
        return SimpleNamespace(nmatvec=4)

    def update_fit(self):
        self.log.append(("fit",))
        return 0.5

    def insert(self, trunc=None, normalize=False, set_ortho=True):
        self.log.append(("insert", trunc, normalize, set_ortho))
        return SimpleNamespace(newdim=7, truncerr=1e-10, decomp=1, jacobi_sweeps=0)

    def maxlinkdim(self):
        return 7


GRAPHS = [ns.path_graph(5), ns.named_comb_tree([2, 3, 1])]


@pytest.mark.parametrize("g", GRAPHS)
@pytest.mark.parametrize("nsites", [1, 2])
def test_eigsolve_hook_sequence_and_per_sweep_truncation(g, nsites):
    net = RecordingNet(g)
    seen_regions, seen_sweeps = [], []

    def region_callback(problem, *, nsweeps, outputlevel, region, region_kwargs, sweep, tag):
        assert nsweeps == 3 and outputlevel == 0 and tag == "user kwarg"
        assert isinstance(problem, ns.EigsolveProblem)
        assert set(region_kwargs) >= {"sweep", "nsites"} or "sweep" in region_kwargs
        seen_regions.append((sweep, tuple(region), problem.eigenvalue))

    def sweep_callback(region_iter, *, nsweeps, outputlevel, sweep, tag):
        seen_sweeps.append((sweep, region_iter.problem.eigenvalue))

    E, state = ns.eigsolve(ns.EigsolveProblem(net=net), nsweeps=3, nsites=nsites,
                           extracter_kwargs=dict(trunc=dict(maxdim=[10, 20]), subspace_algorithm="densitymatrix",
                                                 expansion_factor=1.1),
                           updater_kwargs=dict(krylovdim=4, eager=True),
                           inserter_kwargs=dict(trunc=dict(cutoff=[1e-6, 1e-9], maxdim=[10, 20])),
                           region_callback=region_callback, sweep_callback=sweep_callback, sweep_printer=lambda *a, **k: None,
                           tag="user kwarg")
    plan = [tuple(r) for r, _ in ns.euler_sweep(g, nsites=nsites)]
    assert len(net.log) == 3 * 3 * len(plan)
    for sw in range(3):
        maxdim = [10, 20, 20][sw]            # get_or_last: the last entry repeats (src/truncation_parameters.jl:1-14)
        cutoff = [1e-6, 1e-9, 1e-9][sw]
        for ir, reg in enumerate(plan):
            ex, up, ins = net.log[3 * (sw * len(plan) + ir): 3 * (sw * len(plan) + ir) + 3]
            assert ex[0] == "extract" and ex[1] == reg
            assert ex[2] == (0.0, 1, maxdim)                       # extracter's trunc: cutoff default 0, its own maxdim list
            assert ex[3] == dict(algorithm=L.NSB_EXPAND_DENSITYMATRIX, north_pass=1, expansion_factor=1.1, max_expand=L.INT64_MAX)
            assert up == ("eigsolve", dict(krylovdim=4, maxiter=1, tol=1e-14, which="SR", eager=True))
            assert ins == ("insert", (cutoff, 1, maxdim), False, True)
    assert [(s, r) for s, r, _ in seen_regions] == [(sw, reg) for sw in (1, 2, 3) for reg in plan]
    # the callback sees the problem after the region's update (eigenvalue of that region's solve)
    assert all(e != float("inf") for _, _, e in seen_regions)
    assert [s for s, _ in seen_sweeps] == [1, 2, 3]
    assert E == seen_sweeps[-1][1] == seen_regions[-1][2]
    assert state.net is net


def test_default_truncation_is_unbounded_and_clamped_to_int64():
    net = RecordingNet(ns.path_graph(3))
    ns.eigsolve(ns.EigsolveProblem(net=net), nsweeps=1, nsites=2, sweep_printer=lambda *a, **k: None)
    first = tuple(ns.euler_sweep(net.graph, nsites=2)[0][0])       # the tour starts at default_root_vertex: (3, 2) on this chain
    assert first == (3, 2)
    assert net.log[0] == ("extract", first, (0.0, 1, min(sys.maxsize, L.INT64_MAX)), None)
    assert net.log[2] == ("insert", (0.0, 1, min(sys.maxsize, L.INT64_MAX)), False, True)


@pytest.mark.parametrize("g", GRAPHS)
@pytest.mark.parametrize("nsites", [1, 2])
@pytest.mark.parametrize("order", [1, 2, 4])
def test_tdvp_time_steps_regions_and_next_vertex_follow_the_oracle_plan(g, nsites, order):
    from oracle import region_plans as orp
    from helpers import to_oracle_graph
    net = RecordingNet(g)
    times = [0.1, 0.2, 0.4]
    prob = ns.ApplyExpProblem(net=net)
    exponents = [-1j * t for t in times]
    out = ns.applyexp(prob, exponents, nsites=nsites, tdvp_order=order, sweep_printer=lambda *a, **k: None,
                      updater_kwargs=dict(solver=ns.runge_kutta_solver, order=2),
                      inserter_kwargs=dict(trunc=dict(cutoff=1e-12)))
    assert out.net is net
    # src/applyexp.jl:75: diff([0, exponents...])[2:end] -- the first interval is dropped, one sweep per remaining interval
    steps = [exponents[1] - exponents[0], exponents[2] - exponents[1]]
    og = to_oracle_graph(g)
    expect = []
    for sw, dt in enumerate(steps, start=1):
        plan = orp.tdvp_regions(og, dt, nsites=nsites, tdvp_order=order, sweep=sw, updater_kwargs=dict(order=2))
        for i, (reg, kw) in enumerate(plan):
            nxt = plan[i + 1][0] if i + 1 < len(plan) else None
            expect.append((tuple(reg), kw["updater_kwargs"]["time_step"], nxt, kw.get("nsites", nsites)))
    calls = [net.log[i:i + 3] for i in range(0, len(net.log), 3)]
    assert len(calls) == len(expect)
    t_total = 0.0
    for (ex, up, ins), (reg, dt, nxt, ns_region) in zip(calls, expect):
        assert ex[0] == "extract" and ex[1] == reg
        assert up[0] == "exp" and abs(up[1] - dt) < 1e-15
        assert up[2]["solver"] == "rk" and up[2]["order"] == 2 and up[2]["nsites"] == ns_region
        if ns_region == 1 and nxt is not None and tuple(nxt) != reg:
            path = ns.graphs.vertex_path(g, reg[0], nxt[0])
            assert up[2]["next_vertex"] == path[1]                  # first hop toward the next region (src/applyexp.jl:30-33)
        else:
            assert up[2]["next_vertex"] is None
        assert ins == ("insert", (1e-12, 1, min(sys.maxsize, L.INT64_MAX)), False, True)
        t_total += dt


def test_current_time_accumulates_every_region_step():
    """src/applyexp.jl:44-45: current_time advances by each region's time_step (forward and backward sub-steps cancel so
    that one sweep advances by the sweep's time step)."""
    net = RecordingNet(ns.path_graph(4))
    seen = []
    ns.applyexp(ns.ApplyExpProblem(net=net), [-0.1j, -0.2j, -0.3j], nsites=2, tdvp_order=2,
                sweep_callback=lambda ri, **k: seen.append(ri.problem.current_time), sweep_printer=lambda *a, **k: None)
    assert len(seen) == 2
    assert abs(seen[0] - (-0.1j)) < 1e-14 and abs(seen[1] - (-0.2j)) < 1e-14
    assert ns.process_real_times(seen[1]) == 0.2


def test_krylov_solver_defaults_reach_the_device_call():
    net = RecordingNet(ns.path_graph(3))
    ns.applyexp(ns.ApplyExpProblem(net=net), [-0.1j, -0.2j], nsites=2, tdvp_order=1, sweep_printer=lambda *a, **k: None,
                updater_kwargs=dict(solver=ns.exponentiate_solver, tol=1e-10))
    up = [c for c in net.log if c[0] == "exp"][0]
    assert up[2] == dict(solver="krylov", krylovdim=30, maxiter=100, tol=1e-10, eager=True, nsites=2, next_vertex=None)


def test_error_behaviour_matches_the_reference():
    g = ns.path_graph(4)
    # src/subspace/subspace.jl: no method for this combination of algorithm and problem type
    with pytest.raises(ValueError, match="Subspace expansion"):
        ns.eigsolve(ns.EigsolveProblem(net=RecordingNet(g)), nsweeps=1, nsites=2, extracter_kwargs=dict(subspace_algorithm="nope"))
    with pytest.raises(ValueError, match="Subspace expansion"):
        ns.applyexp(ns.ApplyExpProblem(net=RecordingNet(g)), [-0.1j, -0.2j], nsites=2, extracter_kwargs=dict(subspace_algorithm="ortho"))
    # src/local_solvers/runge_kutta.jl:22: order other than 2 / 4
    with pytest.raises(ValueError, match="must specify `order`"):
        ns.applyexp(ns.ApplyExpProblem(net=RecordingNet(g)), [-0.1j, -0.2j], nsites=2, updater_kwargs=dict(order=3))
    # a solver the device does not run is refused, not silently replaced
    with pytest.raises(TypeError):
        ns.eigsolve(ns.EigsolveProblem(net=RecordingNet(g)), nsweeps=1, nsites=2, updater_kwargs=dict(solver=lambda *a, **k: None))
    # src/inserter.jl:26: regions of other lengths
    prob = ns.EigsolveProblem(net=RecordingNet(g))
    ri = S.RegionIterator(prob, [([1, 2, 3], {})])
    with pytest.raises(ValueError, match="Region of length 3 not currently supported"):
        S.inserter(prob, None, ri, sweep=1)


def test_fitting_problem_never_expands_and_keeps_the_gauge_flagless_insert():
    """src/fitting.jl:38 (expansion commented out) and :79 (set_orthogonal_region=false, normalize passed through)."""
    net = RecordingNet(ns.path_graph(4))
    prob = ns.FittingProblem(net=net)
    it = ns.sweep_iterator(prob, 2, nsites=1, outputlevel=0, extracter_kwargs=dict(subspace_algorithm="densitymatrix"),
                           updater_kwargs={}, inserter_kwargs=dict(normalize=True, set_orthogonal_region=False))
    conv = ns.sweep_solve(it)
    assert conv.overlap == 0.5
    assert all(c[3] is None for c in net.log if c[0] == "extract")
    assert all(c[2:] == (True, False) for c in net.log if c[0] == "insert")
    assert sum(c[0] == "fit" for c in net.log) == 2 * len(ns.euler_sweep(net.graph, nsites=1))


def test_region_iterator_navigation():
    prob = ns.EigsolveProblem(net=RecordingNet(ns.path_graph(3)))
    ri = S.region_iterator(prob, nsites=2, sweep=1, outputlevel=0)      # (the updater requires outputlevel, as in the reference)
    assert S.previous_region(ri) is None and S.current_region(ri) == ri.region_plan[0][0]
    seen = []
    for r in ri:
        seen.append((S.previous_region(r), S.current_region(r), S.next_region(r), S.is_last_region(r)))
    regs = [r for r, _ in ri.region_plan]
    assert [s[1] for s in seen] == regs
    assert [s[0] for s in seen] == [None] + regs[:-1]
    assert [s[2] for s in seen] == regs[1:] + [None]
    assert [s[3] for s in seen] == [False] * (len(regs) - 1) + [True]
