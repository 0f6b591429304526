"""networksolvers_b200 -- B200-native sweep engine behind the NetworkSolvers.jl solver API.

The hot path (H_eff matvec, local Krylov / RK solvers, environment updates, truncating factorisation,
density-matrix subspace expansion) runs in hand-written sm_100a CUDA kernels inside libnsb200.so
(include/nsb200.h); this package is the thin host side: region plans, iterators, kwarg routing and the
three hooks that forward to the C ABI.  There is no CPU fallback.
"""
from ._lib import NsbError, LIB_PATH  # noqa: F401
from .graphs import NamedGraph, path_graph, named_comb_tree, star_of_chains, default_root_vertex  # noqa: F401
from .models import (SiteType, SiteSet, siteinds, OpSum, heisenberg, transverse_ising, hubbard, HostTTN, product_state,  # noqa: F401
                     random_state, ttno, mpo, identity_operator, delta_state, random_tensornetwork, GraphSites, product_operator_sum,
                     ttno_general, compress_operator, operator_direct_sum, expect, inner)
from .device import Context, DeviceNetwork, default_context  # noqa: F401
from .region_plans import (euler_tour_edges, euler_tour_vertices, euler_sweep, post_order_dfs_plan,  # noqa: F401
                           post_order_dfs_sweep, tdvp_sub_time_steps, first_order_sweep, tdvp_regions)
from .solvers import (EigsolveProblem, ApplyExpProblem, FittingProblem, fit_tensornetwork, truncate, apply, overlap, eigenvalue, state, operator, current_time, extracter,  # noqa: F401
                      updater, inserter, RegionIterator, SweepIterator, region_iterator, region_iterator_action,
                      region_plan, region_tuples, sweep_iterator, sweep_solve, eigsolve, dmrg, applyexp, tdvp,
                      eigsolve_solver, exponentiate_solver, runge_kutta_solver, truncation_parameters, get_or_last,
                      compute_expansion, current_region, next_region, previous_region, is_last_region, problem,
                      eigsolve_sweep_printer, applyexp_sweep_printer, default_sweep_printer, process_real_times)
