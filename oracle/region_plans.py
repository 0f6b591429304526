"""Region plans (oracle; test-only).  Restates src/region_plans/{euler_tour,euler_plans,dfs_plans,
tdvp_region_plans}.jl.  A plan is a list of (region, kwargs-dict)."""
from __future__ import annotations

from .graph import default_root_vertex, post_order_dfs_edges, post_order_dfs_vertices


def euler_tour_edges(g, start_vertex):
    """src/region_plans/euler_tour.jl:13-42."""
    visited = set()
    tour = []
    stack = [start_vertex]
    while stack:
        u = stack[-1]
        pushed = False
        for v in g.neighbors(u):
            if (u, v) not in visited:
                visited.add((u, v))
                visited.add((v, u))
                tour.append((u, v))
                stack.append(v)
                pushed = True
                break
        if not pushed:
            stack.pop()
            if stack:
                tour.append((u, stack[-1]))
    return tour


def euler_tour_vertices(g, start_vertex):
    """src/region_plans/euler_tour.jl:44-48."""
    edges = euler_tour_edges(g, start_vertex)
    if not edges:
        return []
    return [edges[0][0]] + [e[1] for e in edges]


def euler_sweep(g, *, nsites, root_vertex=None, **sweep_kwargs):
    """src/region_plans/euler_plans.jl:4-13."""
    if root_vertex is None:
        root_vertex = default_root_vertex(g)
    if nsites == 1:
        return [([v], dict(sweep_kwargs)) for v in euler_tour_vertices(g, root_vertex)]
    elif nsites == 2:
        return [([a, b], dict(sweep_kwargs)) for (a, b) in euler_tour_edges(g, root_vertex)]
    raise ValueError(nsites)


def post_order_dfs_plan(g, *, nsites, root_vertex=None, **sweep_kwargs):
    """src/region_plans/dfs_plans.jl:5-16."""
    if root_vertex is None:
        root_vertex = default_root_vertex(g)
    if nsites == 1:
        return [([v], dict(sweep_kwargs)) for v in post_order_dfs_vertices(g, root_vertex)]
    elif nsites == 2:
        return [([a, b], dict(sweep_kwargs)) for (a, b) in post_order_dfs_edges(g, root_vertex)]
    raise ValueError(nsites)


def post_order_dfs_sweep(g, **kws):
    """src/region_plans/dfs_plans.jl:18-22."""
    fwd = post_order_dfs_plan(g, **kws)
    rev = [(list(reversed(r)), k) for (r, k) in reversed(fwd)]
    return fwd + rev


def tdvp_sub_time_steps(tdvp_order):
    """src/region_plans/tdvp_region_plans.jl:1-12."""
    if tdvp_order == 1:
        return [1.0]
    elif tdvp_order == 2:
        return [0.5, 0.5]
    elif tdvp_order == 4:
        s = 1.0 / (2.0 - 2.0 ** (1.0 / 3.0))
        return [s / 2, s / 2, 0.5 - s, 0.5 - s, s / 2, s / 2]
    raise ValueError(f"TDVP order of {tdvp_order} not supported")


def first_order_sweep(g, time_step, reverse=False, *, updater_kwargs, nsites, **kws):
    """src/region_plans/tdvp_region_plans.jl:14-32."""
    basic = post_order_dfs_plan(g, nsites=nsites, **kws)
    updater_kwargs = {**dict(nsites=nsites, time_step=time_step), **updater_kwargs}
    sweep = []
    for j, (region, region_kws) in enumerate(basic, start=1):
        sweep.append((region, {**dict(nsites=nsites, updater_kwargs=updater_kwargs), **region_kws}))
        if len(region) == 2 and j < len(basic):
            rev_kwargs = dict(updater_kwargs)
            rev_kwargs["time_step"] = -updater_kwargs["time_step"]
            sweep.append(([region[-1]], {**dict(updater_kwargs=rev_kwargs), **region_kws}))
    if reverse:
        sweep = [(list(reversed(r)), k) for (r, k) in reversed(sweep)]
    return sweep


def tdvp_regions(g, time_step, *, updater_kwargs, tdvp_order, nsites, **kws):
    """src/region_plans/tdvp_region_plans.jl:34-44."""
    plan = []
    for step, weight in enumerate(tdvp_sub_time_steps(tdvp_order), start=1):
        plan += first_order_sweep(g, weight * time_step, reverse=(step % 2 == 0),
                                  updater_kwargs=updater_kwargs, nsites=nsites, **kws)
    return plan
