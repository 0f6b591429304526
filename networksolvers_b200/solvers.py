"""The NetworkSolvers.jl solver layer on top of the device library.

Same names, same argument meaning, same error behaviour as the reference so that callers (and the
parity tests) read like the reference's own code:

    problem types      EigsolveProblem, ApplyExpProblem            src/eigsolve.jl:4-12, src/applyexp.jl:4-12
    hooks              extracter, updater, inserter                src/extracter.jl, src/eigsolve.jl:14-28,
                                                                   src/applyexp.jl:18-48, src/inserter.jl
    iterators          RegionIterator, SweepIterator, region_iterator_action, sweep_iterator, region_tuples
                                                                   src/iterators.jl, src/adapters.jl
    drivers            sweep_solve, eigsolve / dmrg, applyexp, tdvp src/sweep_solve.jl, src/eigsolve.jl:45-76,
                                                                   src/applyexp.jl:62-103
    local solvers      eigsolve_solver, exponentiate_solver, runge_kutta_solver (selectable through
                       updater_kwargs=(; solver=...)), executed on the device

The iterator API, region plans and kwarg packs are host-side control flow; every tensor operation is one of
three C-ABI calls per region (nsb_extract / nsb_update_* / nsb_insert).  The state and the projected
operator never leave HBM; `state(problem)` gives a handle whose `.to_host()` downloads on demand.
"""
from __future__ import annotations

import sys

import numpy as np

from . import _lib as L
from . import region_plans as rp
from .device import DeviceNetwork, default_context
from .graphs import vertex_path
from .models import HostTTN

# ---- truncation parameters (src/truncation_parameters.jl) -------------------------------------------
default_maxdim = lambda: sys.maxsize
default_mindim = lambda: 1
default_cutoff = lambda: 0.0


def get_or_last(x, i):
    if isinstance(x, (list, tuple, np.ndarray)):
        return x[-1] if i >= len(x) else x[i - 1]
    return x


def truncation_parameters(sweep, *, cutoff=None, maxdim=None, mindim=None):
    cutoff = default_cutoff() if cutoff is None else cutoff
    maxdim = default_maxdim() if maxdim is None else maxdim
    mindim = default_mindim() if mindim is None else mindim
    return dict(cutoff=get_or_last(cutoff, sweep), mindim=get_or_last(mindim, sweep), maxdim=get_or_last(maxdim, sweep))


def _trunc_tuple(tr):
    return (float(tr["cutoff"]), int(tr["mindim"]), int(min(tr["maxdim"], L.INT64_MAX)))


# ---- subspace expansion parameters (src/subspace/subspace.jl:5-48) ----------------------------------
default_expansion_factor = lambda: 1.5
default_max_expand = lambda: sys.maxsize


def compute_expansion(current_dim, basis_size, *, expansion_factor=1.5, max_expand=sys.maxsize, maxdim=sys.maxsize):
    import math
    e = math.ceil(expansion_factor * current_dim)
    e = min(max_expand, e)
    e = min(basis_size - current_dim, e)
    e = min(maxdim - current_dim, e)
    return max(0, e)


# ---- local solver descriptors ------------------------------------------------------------------------
class _DeviceSolver:
    """A local solver that runs on the device; passed as `updater_kwargs=dict(solver=...)` exactly like
    the reference's solver functions (src/local_solvers/*.jl)."""

    def __init__(self, name, kind, defaults):
        self.__name__ = name
        self.kind = kind
        self.defaults = defaults

    def __repr__(self):
        return f"<device solver {self.__name__}>"


eigsolve_solver = _DeviceSolver("eigsolve_solver", "eigsolve",
                                dict(which_eigval="SR", ishermitian=True, tol=1e-14, krylovdim=3, maxiter=1,
                                     verbosity=0, eager=False))
exponentiate_solver = _DeviceSolver("exponentiate_solver", "krylov",
                                    dict(krylovdim=30, maxiter=100, verbosity=0, tol=1e-12, ishermitian=True,
                                         issymmetric=True, eager=True))
runge_kutta_solver = _DeviceSolver("runge_kutta_solver", "rk", dict(order=4))


# ---- problems ---------------------------------------------------------------------------------------
class DeviceState:
    """Handle returned by `state(problem)`; downloads lazily."""

    def __init__(self, net: DeviceNetwork):
        self.net = net
        self.graph = net.graph

    def to_host(self) -> HostTTN:
        return self.net.to_host()

    def maxlinkdim(self):
        return self.net.maxlinkdim()

    def linkdims(self):
        return self.net.linkdims()

    def norm(self):
        return self.net.norm()


class _Problem:
    def setproperties(self, **kw):
        new = self.__class__.__new__(self.__class__)
        new.__dict__.update(self.__dict__)
        new.__dict__.update(kw)
        return new


class EigsolveProblem(_Problem):
    def __init__(self, state=None, operator=None, eigenvalue=float("inf"), *, net=None, ctx=None, dtype=None):
        self.net = net if net is not None else DeviceNetwork(operator, state, dtype=dtype, ctx=ctx)
        self.eigenvalue = eigenvalue
        self.last_truncerr = 0.0
        self.last_info = {}

    @property
    def state(self):
        return DeviceState(self.net)

    @property
    def operator(self):
        return self.net


class ApplyExpProblem(_Problem):
    def __init__(self, state=None, operator=None, current_time=0.0, *, net=None, ctx=None, dtype=None):
        self.net = net if net is not None else DeviceNetwork(operator, state, dtype=dtype or np.complex128, ctx=ctx)
        self.current_time = current_time
        self.last_truncerr = 0.0
        self.last_info = {}

    @property
    def state(self):
        return DeviceState(self.net)

    @property
    def operator(self):
        return self.net


class FittingProblem(_Problem):
    """src/fitting.jl:8-18.  The ket being fitted is the network's state; the target |x> and the operator A of the
    overlap network <psi| A |x> are resident on the device (DeviceNetwork.set_fit_target)."""

    def __init__(self, state=None, target=None, operator=None, overlap=0.0, *, net=None, ctx=None, dtype=None):
        if net is None:
            net = DeviceNetwork(operator, state, dtype=dtype, ctx=ctx)
            net.set_fit_target(target)
        self.net = net
        self.overlap = overlap
        self.last_truncerr = 0.0
        self.last_info = {}

    @property
    def state(self):
        return DeviceState(self.net)


overlap = lambda F: F.overlap
eigenvalue = lambda E: E.eigenvalue
state = lambda P: P.state
operator = lambda P: P.operator
current_time = lambda T: T.current_time


class LocalState:
    """Token for the local tensor, which stays on the device between extracter / updater / inserter."""

    def __init__(self, net):
        self.net = net

    def array(self):
        return self.net.local_download()


# ---- hooks ------------------------------------------------------------------------------------------
def extracter(problem, region_iterator, *, sweep, trunc=None, subspace_algorithm=None, north_pass=1,
              expansion_factor=None, max_expand=None, **kws):
    trunc = truncation_parameters(sweep, **(trunc or {}))
    region = current_region(region_iterator)
    expand = None
    if subspace_algorithm is not None:
        algs = {"densitymatrix": L.NSB_EXPAND_DENSITYMATRIX, "ortho": L.NSB_EXPAND_ORTHO}
        if subspace_algorithm not in algs:
            raise ValueError("Subspace expansion (subspace_expand!) not defined for requested combination of "
                             "subspace_algorithm and problem types")
        # "ortho" (src/subspace/ortho_subspace.jl:19-77) is only defined for EigsolveProblem in the reference
        if subspace_algorithm == "ortho" and not isinstance(problem, EigsolveProblem):
            raise ValueError("Subspace expansion (subspace_expand!) not defined for requested combination of "
                             "subspace_algorithm and problem types")
        expand = dict(algorithm=algs[subspace_algorithm], north_pass=north_pass,
                      expansion_factor=default_expansion_factor() if expansion_factor is None else expansion_factor,
                      max_expand=min(default_max_expand() if max_expand is None else max_expand, L.INT64_MAX))
    if isinstance(problem, FittingProblem):
        expand = None        # src/fitting.jl:38: the expansion call is commented out in the reference
    info = problem.net.extract(region, _trunc_tuple(trunc), expand)
    problem.last_info = dict(expanded=info.expanded, env_builds=info.env_builds, qr_steps=info.qr_steps)
    return problem, LocalState(problem.net)


def updater(problem, local_state, region_iterator, **kws):
    if isinstance(problem, EigsolveProblem):
        return _updater_eigsolve(problem, local_state, region_iterator, **kws)
    if isinstance(problem, ApplyExpProblem):
        return _updater_applyexp(problem, local_state, region_iterator, **kws)
    if isinstance(problem, FittingProblem):
        return _updater_fitting(problem, local_state, region_iterator, **kws)
    raise TypeError(f"no updater for {type(problem).__name__}")


def _updater_eigsolve(E, local_state, region_iterator, *, outputlevel, solver=eigsolve_solver, **kws):
    if not isinstance(solver, _DeviceSolver) or solver.kind != "eigsolve":
        raise TypeError("EigsolveProblem on the device needs solver=eigsolve_solver")
    p = {**solver.defaults, **kws}
    eigval, info = E.net.update_eigsolve(krylovdim=p["krylovdim"], maxiter=p["maxiter"], tol=p["tol"],
                                         which=p["which_eigval"], eager=p["eager"])
    E = E.setproperties(eigenvalue=eigval)
    E.last_info = dict(E.last_info, nmatvec=info.nmatvec, residual=info.residual)
    if outputlevel >= 2:
        print("  Region %s: energy = %.12f" % (current_region(region_iterator), E.eigenvalue))
    return E, local_state


def _updater_applyexp(T, local_state, region_iterator, *, nsites, time_step, solver=runge_kutta_solver,
                      outputlevel, **kws):
    if not isinstance(solver, _DeviceSolver) or solver.kind not in ("rk", "krylov"):
        raise TypeError("ApplyExpProblem on the device needs solver=runge_kutta_solver or exponentiate_solver")
    p = {**solver.defaults, **kws}
    next_vertex = None
    if nsites == 1:
        curr_reg, next_reg = current_region(region_iterator), next_region(region_iterator)
        if next_reg is not None and next_reg != curr_reg:
            next_vertex = vertex_path(T.net.graph, curr_reg[0], next_reg[0])[1]
    if solver.kind == "rk":
        if p.get("order") not in (2, 4):
            raise ValueError("For runge_kutta_solver, must specify `order` keyword")
        info = T.net.update_exp(time_step, solver="rk", order=p["order"], nsites=nsites, next_vertex=next_vertex)
    else:
        info = T.net.update_exp(time_step, solver="krylov", krylovdim=p["krylovdim"], maxiter=p["maxiter"],
                                tol=p["tol"], eager=p["eager"], nsites=nsites, next_vertex=next_vertex)
    T = T.setproperties(current_time=T.current_time + time_step)
    T.last_info = dict(T.last_info, nmatvec=info.nmatvec)
    return T, local_state


def _updater_fitting(F, local_state, region_iterator, *, outputlevel, **kws):
    """src/fitting.jl:42-49: overlap = n / sqrt(n), n = <local|local>."""
    F = F.setproperties(overlap=F.net.update_fit())
    if outputlevel >= 2:
        print("  Region %s: squared overlap = %.12f" % (current_region(region_iterator), F.overlap))
    return F, local_state


def inserter(problem, local_tensor, region_iterator, *, normalize=False, set_orthogonal_region=True, sweep,
             trunc=None, **kws):
    trunc = truncation_parameters(sweep, **(trunc or {}))
    region = current_region(region_iterator)
    if len(region) not in (1, 2):
        raise ValueError(f"Region of length {len(region)} not currently supported")
    info = problem.net.insert(_trunc_tuple(trunc), normalize, set_orthogonal_region)
    problem.last_truncerr = info.truncerr
    problem.last_info = dict(problem.last_info, newdim=info.newdim, truncerr=info.truncerr, decomp=info.decomp,
                             jacobi_sweeps=info.jacobi_sweeps)
    return problem


# ---- iterators (src/iterators.jl, src/adapters.jl) -----------------------------------------------------
class RegionIterator:
    def __init__(self, problem, region_plan, which_region=1):
        self.problem, self.region_plan, self.which_region = problem, region_plan, which_region

    @property
    def state(self):      # convenience for sweep callbacks written against `problem.state`
        return self.problem.state

    def __iter__(self):
        for which in range(1, len(self.region_plan) + 1):
            self.which_region = which
            _, kwargs = self.region_plan[which - 1]
            # the reference dispatches `region_iterator_action!` on the problem type (examples/timed_dmrg/timed_eigsolve.jl:41-75
            # wraps an EigsolveProblem and times the hooks): a problem object may bring its own action as a method
            action = getattr(self.problem, "region_iterator_action", None)
            self.problem = action(self, **kwargs) if action is not None else region_iterator_action(self.problem, self, **kwargs)
            yield self


problem = lambda R: R.problem() if isinstance(R, SweepIterator) else R.problem
current_region_plan = lambda R: R.region_plan[R.which_region - 1]
current_region = lambda R: current_region_plan(R)[0]
region_kwargs = lambda R: current_region_plan(R)[1]
previous_region = lambda R: None if R.which_region == 1 else R.region_plan[R.which_region - 2][0]
next_region = lambda R: None if R.which_region == len(R.region_plan) else R.region_plan[R.which_region][0]
is_last_region = lambda R: next_region(R) is None


def region_plan(problem, **kws):
    if isinstance(problem, ApplyExpProblem):
        kws = dict(kws)
        nsites, time_step = kws.pop("nsites"), kws.pop("time_step")
        return rp.tdvp_regions(problem.net.graph, time_step, nsites=nsites, **kws)
    return rp.euler_sweep(problem.net.graph, **kws)


def region_iterator(problem, **sweep_kwargs):
    return RegionIterator(problem, region_plan(problem, **sweep_kwargs))


def region_iterator_action(problem, region_iterator, *, extracter_kwargs=None, updater_kwargs=None,
                           inserter_kwargs=None, sweep, **kws):
    problem, local_state = extracter(problem, region_iterator, **{**(extracter_kwargs or {}), "sweep": sweep, **kws})
    problem, local_state = updater(problem, local_state, region_iterator, **{**(updater_kwargs or {}), **kws})
    problem = inserter(problem, local_state, region_iterator, **{"sweep": sweep, **(inserter_kwargs or {}), **kws})
    return problem


def region_tuples(R):
    """Adapter: iterate (current_region, region_kwargs) tuples (src/adapters.jl:12-32)."""
    for it in R:
        yield current_region_plan(it)


class SweepIterator:
    def __init__(self, problem, sweep_kws):
        self.sweep_kws = list(sweep_kws)
        self.region_iter = region_iterator(problem, sweep=1, **self.sweep_kws[0])
        self.which_sweep = 1

    def problem(self):
        return self.region_iter.problem

    def __len__(self):
        return len(self.sweep_kws)

    def __iter__(self):
        for i, kws in enumerate(self.sweep_kws):
            if i > 0:
                self.region_iter = region_iterator(self.region_iter.problem, sweep=self.which_sweep, **kws)
            self.which_sweep += 1
            yield self.region_iter


def sweep_iterator(problem, sweep_kws_or_nsweeps, **sweep_kws):
    if isinstance(sweep_kws_or_nsweeps, int):
        return SweepIterator(problem, [dict(sweep_kws) for _ in range(sweep_kws_or_nsweeps)])
    return SweepIterator(problem, sweep_kws_or_nsweeps)


# ---- drivers (src/sweep_solve.jl, src/eigsolve.jl, src/applyexp.jl) -------------------------------------
def default_region_callback(problem, **kws):
    return None


def default_sweep_callback(problem, **kws):
    return None


def default_sweep_printer(problem, *, outputlevel, sweep, nsweeps, **kws):
    if outputlevel >= 1:
        print(f"Done with sweep {sweep}/{nsweeps}")


def sweep_solve(sweep_iterator, *, outputlevel=0, region_callback=default_region_callback,
                sweep_callback=default_sweep_callback, sweep_printer=default_sweep_printer, **kwargs):
    nsweeps = len(sweep_iterator)
    for sweep, region_iter in enumerate(sweep_iterator, start=1):
        for region, region_kwargs in region_tuples(region_iter):
            region_callback(region_iter.problem, nsweeps=nsweeps, outputlevel=outputlevel, region=region,
                            region_kwargs=region_kwargs, sweep=sweep, **kwargs)
        sweep_callback(region_iter, nsweeps=nsweeps, outputlevel=outputlevel, sweep=sweep, **kwargs)
        sweep_printer(region_iter, nsweeps=nsweeps, outputlevel=outputlevel, sweep=sweep, **kwargs)
    return sweep_iterator.problem()


def eigsolve_sweep_printer(region_iterator, *, outputlevel, sweep, nsweeps, **kws):
    if outputlevel >= 1:
        E = region_iterator.problem
        fmt = "After sweep %02d/%d " if nsweeps >= 10 else "After sweep %d/%d "
        print(fmt % (sweep, nsweeps) + "eigenvalue=%.12f maxlinkdim=%d" % (E.eigenvalue, E.net.maxlinkdim()), flush=True)


def eigsolve(*args, nsweeps, nsites=1, outputlevel=0, extracter_kwargs=None, updater_kwargs=None,
             inserter_kwargs=None, sweep_printer=eigsolve_sweep_printer, ctx=None, **kws):
    """eigsolve(H, init_state; ...) or eigsolve(init_prob; ...)  ->  (eigenvalue, state)."""
    if len(args) == 2:
        H, init_state = args
        init_prob = EigsolveProblem(state=init_state, operator=H, ctx=ctx)
    else:
        (init_prob,) = args
    sweep_iter = sweep_iterator(init_prob, nsweeps, nsites=nsites, outputlevel=outputlevel,
                                extracter_kwargs=extracter_kwargs or {}, updater_kwargs=updater_kwargs or {},
                                inserter_kwargs=inserter_kwargs or {})
    prob = sweep_solve(sweep_iter, outputlevel=outputlevel, sweep_printer=sweep_printer, **kws)
    return prob.eigenvalue, prob.state


def dmrg(*args, **kws):
    return eigsolve(*args, **kws)


def applyexp_sweep_printer(region_iterator, *, outputlevel, sweep, nsweeps, process_time=lambda z: z, **kws):
    if outputlevel >= 1:
        T = region_iterator.problem
        print("  Current time = %s, maxlinkdim=%d" % (process_time(T.current_time), T.net.maxlinkdim()), flush=True)


def applyexp(*args, extracter_kwargs=None, updater_kwargs=None, inserter_kwargs=None, outputlevel=0, nsites=1,
             tdvp_order=4, sweep_printer=applyexp_sweep_printer, ctx=None, **kws):
    """applyexp(H, init_state, exponents; ...) or applyexp(init_prob, exponents; ...)  ->  state."""
    if len(args) == 3:
        H, init_state, exponents = args
        init_prob = ApplyExpProblem(state=init_state, operator=H, ctx=ctx)
    else:
        init_prob, exponents = args
    ex = [0.0] + list(exponents)
    time_steps = [ex[i + 1] - ex[i] for i in range(len(ex) - 1)][1:]
    sweep_kws = dict(outputlevel=outputlevel, extracter_kwargs=extracter_kwargs or {}, inserter_kwargs=inserter_kwargs or {},
                     nsites=nsites, tdvp_order=tdvp_order, updater_kwargs=updater_kwargs or {})
    kws_array = [dict(sweep_kws, time_step=t) for t in time_steps]
    sweep_iter = sweep_iterator(init_prob, kws_array)
    prob = sweep_solve(sweep_iter, outputlevel=outputlevel, sweep_printer=sweep_printer, **kws)
    return prob.state


# ---- fitting (src/fitting.jl:55-112) ---------------------------------------------------------------------------
def fit_tensornetwork(target, operator, init_state, *, nsweeps=25, nsites=1, outputlevel=0, extracter_kwargs=None,
                      updater_kwargs=None, inserter_kwargs=None, normalize=True, ctx=None, dtype=None, **kws):
    """Fit `init_state` to operator |target> by sweeping (src/fitting.jl:55-84).  The reference builds the overlap
    network with `itn.inner_network`; here its three layers are passed separately (operator=None: identity)."""
    from .models import identity_operator, GraphSites
    if dtype is None:
        dtype = np.result_type(target.dtype(), init_state.dtype(), *([operator.dtype()] if operator is not None else []))
        dtype = np.complex128 if np.issubdtype(dtype, np.complexfloating) else np.float64
    if operator is None:
        operator = identity_operator(GraphSites.of(init_state), dtype)
    prob = FittingProblem(state=init_state, target=target, operator=operator, ctx=ctx, dtype=dtype)
    ik = dict(inserter_kwargs or {}, normalize=normalize, set_orthogonal_region=False)
    sweep_iter = sweep_iterator(prob, nsweeps, nsites=nsites, outputlevel=outputlevel,
                                extracter_kwargs=extracter_kwargs or {}, updater_kwargs=updater_kwargs or {}, inserter_kwargs=ik)
    conv = sweep_solve(sweep_iter, outputlevel=outputlevel, **kws)
    return conv.state


def truncate(tn, *, maxdim, cutoff=0.0, **kws):
    """`itn.truncate(tn; maxdim, cutoff)` (src/fitting.jl:90-97): fit a delta-initialised network of link dimension
    maxdim to tn."""
    from .models import delta_state, GraphSites
    init = delta_state(GraphSites.of(tn), maxdim, tn.dtype())
    return fit_tensornetwork(tn, None, init, inserter_kwargs=dict(trunc=dict(cutoff=cutoff, maxdim=maxdim)), **kws)


def apply(A, x, *, maxdim, cutoff=0.0, **kws):
    """`itn.apply(A, x; maxdim, cutoff)` (src/fitting.jl:99-112): fit to A|x>."""
    from .models import delta_state, GraphSites
    init = delta_state(GraphSites.of(x), maxdim, np.result_type(A.dtype(), x.dtype()))
    return fit_tensornetwork(x, A, init, inserter_kwargs=dict(trunc=dict(cutoff=cutoff, maxdim=maxdim)), **kws)


def process_real_times(z):
    return round(-complex(z).imag, 10)


def tdvp(H, init_state, time_points, *, process_time=process_real_times, sweep_printer=None, **kws):
    if sweep_printer is None:
        def sweep_printer(*a, **k):
            return applyexp_sweep_printer(*a, process_time=process_time, **k)
    exponents = [-1j * t for t in time_points]
    return applyexp(H, init_state, exponents, sweep_printer=sweep_printer, **kws)
