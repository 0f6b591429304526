"""Problems, the three region hooks, iterators and drivers (oracle; test-only).

Restates src/eigsolve.jl, src/applyexp.jl, src/extracter.jl, src/inserter.jl, src/iterators.jl,
src/adapters.jl, src/sweep_solve.jl.  Control flow and kwarg routing follow the reference line by
line."""
from __future__ import annotations

import numpy as np

from . import region_plans as rp
from .gauge import orthogonalize
from .graph import tree_path
from .local_solvers import eigsolve_solver, runge_kutta_solver
from .operator_map import optimal_map
from .projttn import ProjTTN, position
from .subspace import subspace_expand
from .tensor import Tensor, contract, factorize, link, qr, uniquelabels
from .truncation_parameters import truncation_parameters


class _Problem:
    def setproperties(self, **kw):
        new = self.__class__.__new__(self.__class__)
        new.__dict__.update(self.__dict__)
        new.__dict__.update(kw)
        return new


class EigsolveProblem(_Problem):
    """src/eigsolve.jl:4-12."""

    def __init__(self, state, operator, eigenvalue=np.inf):
        self.state, self.operator, self.eigenvalue = state, operator, eigenvalue


class ApplyExpProblem(_Problem):
    """src/applyexp.jl:4-12."""

    def __init__(self, state, operator, current_time=0.0):
        self.state, self.operator, self.current_time = state, operator, current_time


# --------------------------------------------------------------------------- region iterator
class RegionIterator:
    """src/iterators.jl:46-71."""

    def __init__(self, problem, region_plan, which_region=1):
        self.problem, self.region_plan, self.which_region = problem, region_plan, which_region

    def current_region_plan(self):
        return self.region_plan[self.which_region - 1]

    def current_region(self):
        return self.current_region_plan()[0]

    def region_kwargs(self):
        return self.current_region_plan()[1]

    def previous_region(self):
        return None if self.which_region == 1 else self.region_plan[self.which_region - 2][0]

    def next_region(self):
        return None if self.which_region == len(self.region_plan) else self.region_plan[self.which_region][0]

    def is_last_region(self):
        return self.next_region() is None

    def __iter__(self):
        for which in range(1, len(self.region_plan) + 1):
            self.which_region = which
            region, kwargs = self.region_plan[which - 1]
            self.problem = region_iterator_action(self.problem, self, **kwargs)
            yield region, kwargs


def region_plan(problem, **kws):
    """src/iterators.jl:102-104 and the ApplyExpProblem override src/applyexp.jl:14-16."""
    if isinstance(problem, ApplyExpProblem):
        kws = dict(kws)
        nsites = kws.pop("nsites")
        time_step = kws.pop("time_step")
        return rp.tdvp_regions(problem.state.graph, time_step, nsites=nsites, **kws)
    if _is_fitting(problem):
        from . import fitting
        return fitting.region_plan(problem, **kws)       # src/fitting.jl:51-53
    return rp.euler_sweep(problem.state.graph, **kws)


def region_iterator(problem, **sweep_kwargs):
    """src/iterators.jl:77-79."""
    return RegionIterator(problem, region_plan(problem, **sweep_kwargs))


def region_iterator_action(problem, region_iter, *, extracter_kwargs=None, updater_kwargs=None,
                           inserter_kwargs=None, sweep, **kws):
    """src/iterators.jl:81-100."""
    extracter_kwargs = extracter_kwargs or {}
    updater_kwargs = updater_kwargs or {}
    inserter_kwargs = inserter_kwargs or {}
    problem, local_state = extracter(problem, region_iter, **{**extracter_kwargs, "sweep": sweep, **kws})
    problem, local_state = updater(problem, local_state, region_iter, **{**updater_kwargs, **kws})
    problem = inserter(problem, local_state, region_iter, **{"sweep": sweep, **inserter_kwargs, **kws})
    return problem


# --------------------------------------------------------------------------- hooks
COUNTERS = {}


def _is_fitting(problem):
    from .fitting import FittingProblem
    return isinstance(problem, FittingProblem)


def extracter(problem, region_iter, *, sweep, trunc=None, **kws):
    """src/extracter.jl:3-17 (and the FittingProblem method, src/fitting.jl:25-40)."""
    if _is_fitting(problem):
        from . import fitting
        return fitting.extracter(problem, region_iter, sweep=sweep, **kws)
    trunc = truncation_parameters(sweep, **(trunc or {}))
    region = region_iter.current_region()
    psi = orthogonalize(problem.state, region)
    local_state = psi[region[0]]
    for v in region[1:]:
        local_state = contract(local_state, psi[v])
    problem = problem.setproperties(state=psi)
    problem, local_state = subspace_expand(problem, local_state, region_iter, sweep=sweep, trunc=trunc, **kws)
    shifted = position(problem.operator, problem.state, region, COUNTERS)
    return problem.setproperties(operator=shifted), local_state


def inserter(problem, local_tensor, region_iter, *, normalize=False, set_orthogonal_region=True,
             sweep, trunc=None, **kws):
    """src/inserter.jl:3-33."""
    trunc = truncation_parameters(sweep, **(trunc or {}))
    region = region_iter.current_region()
    psi = problem.state.copy()
    if len(region) == 1:
        C = local_tensor
    elif len(region) == 2:
        a, b = region
        left = [l for l in psi[a].labels if l in local_tensor.labels]
        if psi.qn is not None:
            U, C, info = _factorize_qn(psi, local_tensor, a, b, left, **trunc)
        else:
            U, C, info = factorize(local_tensor, left, link(a, b), **trunc)
        psi[a] = U
        COUNTERS["last_truncerr"] = info["truncerr"]
        COUNTERS.setdefault("truncerrs", []).append(info["truncerr"])
    else:
        raise ValueError(f"Region of length {len(region)} not currently supported")
    v = region[-1]
    psi[v] = C
    if set_orthogonal_region:
        psi.ortho_region = [v]
    if normalize:
        psi[v] = psi[v] / psi[v].norm()
    return problem.setproperties(state=psi)


def _factorize_qn(psi, theta, a, b, left, *, cutoff, mindim, maxdim):
    """QN-conserving `factorize` (block-wise, merged-spectrum truncation); updates psi.qn for the new bond."""
    from .qn import label_charges, multi_index_charges, svd_trunc_qn
    qn = psi.qn
    right = [l for l in theta.labels if l not in left]
    dl = int(np.prod([theta.dim(l) for l in left]))
    dr = int(np.prod([theta.dim(l) for l in right]))
    M = theta.array(left + right).reshape(dl, dr)
    row_keys = multi_index_charges([label_charges(qn, a, x) for x in left])
    col_keys = qn.total[None, :] - multi_index_charges([label_charges(qn, b, x) for x in right])
    maxdim = min(maxdim, dl, dr)
    use_eigen = cutoff > 1e-12
    Um, spec, Rm, newk, terr = svd_trunc_qn(M, row_keys, col_keys, cutoff=cutoff, mindim=mindim, maxdim=maxdim, use_eigen=use_eigen)
    k = Um.shape[1]
    bond = link(a, b)
    U = Tensor(Um.reshape([theta.dim(l) for l in left] + [k]), left + [bond])
    C = Tensor(Rm.reshape([k] + [theta.dim(l) for l in right]), [bond] + right)
    qn.set_link(a, b, newk)
    return U, C, {"decomp": "eigen" if use_eigen else "svd", "truncerr": terr, "spectrum": spec}


def updater(problem, local_state, region_iter, **kws):
    if _is_fitting(problem):
        from . import fitting
        return fitting.updater(problem, local_state, region_iter, **kws)
    if isinstance(problem, EigsolveProblem):
        return _updater_eigsolve(problem, local_state, region_iter, **kws)
    return _updater_applyexp(problem, local_state, region_iter, **kws)


def _updater_eigsolve(E, local_state, region_iter, *, outputlevel, solver=eigsolve_solver, **kws):
    """src/eigsolve.jl:14-28."""
    def op(x):
        COUNTERS["matvecs"] = COUNTERS.get("matvecs", 0) + 1
        return optimal_map(E.operator, x)

    eigval, local_state = solver(op, local_state, **kws)
    E = E.setproperties(eigenvalue=eigval)
    if outputlevel >= 2:
        print("  Region %s: energy = %.12f" % (region_iter.current_region(), E.eigenvalue))
    return E, local_state


_QR = ("x", "qr", 0)


def _updater_applyexp(T, local_state, region_iter, *, nsites, time_step, solver=runge_kutta_solver,
                      outputlevel, **kws):
    """src/applyexp.jl:18-48."""
    def op_for(P):
        def op(x):
            COUNTERS["matvecs"] = COUNTERS.get("matvecs", 0) + 1
            return optimal_map(P, x)
        return op

    local_state, info = solver(op_for(T.operator), time_step, local_state, **kws)
    if nsites == 1:
        curr_reg = region_iter.current_region()
        next_reg = region_iter.next_region()
        if next_reg is not None and next_reg != curr_reg:
            path = tree_path(T.state.graph, curr_reg[0], next_reg[0])
            v1, v2 = path[0], path[1]
            psi = T.state.copy()
            left = uniquelabels(local_state, psi[v2])
            Q, R = qr(local_state, left, _QR)
            psi[v1] = Q
            shifted = position(T.operator, psi, [("edge", v1, v2)], COUNTERS)
            R_t, _ = solver(op_for(shifted), -time_step, R, **kws)
            local_state = contract(Q, R_t).permute(local_state.labels)
    T = T.setproperties(current_time=T.current_time + time_step)
    return T, local_state


# --------------------------------------------------------------------------- sweeps
class SweepIterator:
    """src/iterators.jl:5-40."""

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


def default_sweep_printer(region_iter, *, outputlevel, sweep, nsweeps, **kws):
    if outputlevel >= 1:
        print(f"Done with sweep {sweep}/{nsweeps}")


def sweep_solve(sweep_iter, *, outputlevel=0, region_callback=None, sweep_callback=None,
                sweep_printer=default_sweep_printer, **kwargs):
    """src/sweep_solve.jl:12-40."""
    nsweeps = len(sweep_iter)
    for sweep, region_iter in enumerate(sweep_iter, start=1):
        for region, region_kwargs in region_iter:
            if region_callback is not None:
                region_callback(region_iter.problem, nsweeps=nsweeps, outputlevel=outputlevel,
                                region=region, region_kwargs=region_kwargs, sweep=sweep, **kwargs)
        if sweep_callback is not None:
            sweep_callback(region_iter, nsweeps=nsweeps, outputlevel=outputlevel, sweep=sweep, **kwargs)
        sweep_printer(region_iter, nsweeps=nsweeps, outputlevel=outputlevel, sweep=sweep, **kwargs)
    return sweep_iter.problem()


def eigsolve_sweep_printer(region_iter, *, outputlevel, sweep, nsweeps, **kws):
    """src/eigsolve.jl:30-43."""
    if outputlevel >= 1:
        E = region_iter.problem
        print("After sweep %d/%d eigenvalue=%.12f maxlinkdim=%d" % (sweep, nsweeps, E.eigenvalue,
                                                                     E.state.maxlinkdim()))


def eigsolve(H, init_state, *, nsweeps, nsites=1, outputlevel=0, extracter_kwargs=None,
             updater_kwargs=None, inserter_kwargs=None, sweep_printer=eigsolve_sweep_printer, **kws):
    """src/eigsolve.jl:45-74."""
    prob = EigsolveProblem(state=init_state, operator=ProjTTN(H))
    it = sweep_iterator(prob, nsweeps, nsites=nsites, outputlevel=outputlevel,
                        extracter_kwargs=extracter_kwargs or {}, updater_kwargs=updater_kwargs or {},
                        inserter_kwargs=inserter_kwargs or {})
    prob = sweep_solve(it, outputlevel=outputlevel, sweep_printer=sweep_printer, **kws)
    return prob.eigenvalue, prob.state


dmrg = eigsolve


def applyexp_sweep_printer(region_iter, *, outputlevel, sweep, nsweeps, process_time=lambda z: z, **kws):
    """src/applyexp.jl:50-60."""
    if outputlevel >= 1:
        T = region_iter.problem
        print("  Current time = %s, maxlinkdim=%d" % (process_time(T.current_time), T.state.maxlinkdim()))


def applyexp(H, init_state, exponents, *, extracter_kwargs=None, updater_kwargs=None,
             inserter_kwargs=None, outputlevel=0, nsites=1, tdvp_order=4,
             sweep_printer=applyexp_sweep_printer, **kws):
    """src/applyexp.jl:62-89."""
    prob = ApplyExpProblem(state=init_state, operator=ProjTTN(H))
    ex = [0.0] + list(exponents)
    time_steps = [ex[i + 1] - ex[i] for i in range(len(ex) - 1)][1:]
    sweep_kws = dict(outputlevel=outputlevel, extracter_kwargs=extracter_kwargs or {},
                     inserter_kwargs=inserter_kwargs or {}, nsites=nsites, tdvp_order=tdvp_order,
                     updater_kwargs=updater_kwargs or {})
    kws_array = [dict(sweep_kws, time_step=t) for t in time_steps]
    it = sweep_iterator(prob, kws_array)
    prob = sweep_solve(it, outputlevel=outputlevel, sweep_printer=sweep_printer, **kws)
    return prob.state


def process_real_times(z):
    return round(-complex(z).imag, 10)


def tdvp(H, init_state, time_points, *, process_time=process_real_times, sweep_printer=None, **kws):
    """src/applyexp.jl:93-103."""
    if sweep_printer is None:
        def sweep_printer(*a, **k):
            return applyexp_sweep_printer(*a, process_time=process_time, **k)
    exponents = [-1j * t for t in time_points]
    return applyexp(H, init_state, exponents, sweep_printer=sweep_printer, **kws)
