"""Region plans (host side; pure control flow, no tensor data).  Mirrors the reference's
src/region_plans/euler_tour.jl, euler_plans.jl, dfs_plans.jl and tdvp_region_plans.jl: same function
names, same kwarg routing; a plan is a list of (region, kwargs) pairs."""
from __future__ import annotations

from .graphs import default_root_vertex, post_order_dfs_edges, post_order_dfs_vertices


def euler_tour_edges(graph, start_vertex):
    seen = set()
    tour, stack = [], [start_vertex]
    while stack:
        u = stack[-1]
        nxt = next((v for v in graph.neighbors(u) if (u, v) not in seen), None)
        if nxt is not None:
            seen.update({(u, nxt), (nxt, u)})
            tour.append((u, nxt))
            stack.append(nxt)
        else:
            stack.pop()
            if stack:
                tour.append((u, stack[-1]))
    return tour


def euler_tour_vertices(graph, start_vertex):
    edges = euler_tour_edges(graph, start_vertex)
    return [edges[0][0]] + [b for _, b in edges] if edges else []


def euler_sweep(graph, *, nsites, root_vertex=None, **sweep_kwargs):
    root = default_root_vertex(graph) if root_vertex is None else root_vertex
    if nsites == 1:
        return [([v], dict(sweep_kwargs)) for v in euler_tour_vertices(graph, root)]
    if nsites == 2:
        return [([a, b], dict(sweep_kwargs)) for a, b in euler_tour_edges(graph, root)]
    raise ValueError(f"nsites={nsites} not supported")


def post_order_dfs_plan(graph, *, nsites, root_vertex=None, **sweep_kwargs):
    root = default_root_vertex(graph) if root_vertex is None else root_vertex
    if nsites == 1:
        return [([v], dict(sweep_kwargs)) for v in post_order_dfs_vertices(graph, root)]
    if nsites == 2:
        return [([a, b], dict(sweep_kwargs)) for a, b in post_order_dfs_edges(graph, root)]
    raise ValueError(f"nsites={nsites} not supported")


def post_order_dfs_sweep(graph, **kws):
    fwd = post_order_dfs_plan(graph, **kws)
    return fwd + [(r[::-1], k) for r, k in fwd[::-1]]


def tdvp_sub_time_steps(tdvp_order):
    if tdvp_order == 1:
        return [1.0]
    if tdvp_order == 2:
        return [1 / 2, 1 / 2]
    if tdvp_order == 4:
        s = (2 - 2 ** (1 / 3)) ** (-1)
        return [s / 2, s / 2, 1 / 2 - s, 1 / 2 - s, s / 2, s / 2]
    raise ValueError(f"TDVP order of {tdvp_order} not supported")


def first_order_sweep(graph, time_step, reverse=False, *, updater_kwargs, nsites, **kws):
    basic = post_order_dfs_plan(graph, nsites=nsites, **kws)
    upd = {"nsites": nsites, "time_step": time_step, **updater_kwargs}
    sweep = []
    for j, (region, region_kws) in enumerate(basic):
        sweep.append((region, {"nsites": nsites, "updater_kwargs": upd, **region_kws}))
        if len(region) == 2 and j < len(basic) - 1:
            back = {**upd, "time_step": -upd["time_step"]}
            sweep.append(([region[-1]], {"updater_kwargs": back, **region_kws}))
    if reverse:
        sweep = [(r[::-1], k) for r, k in sweep[::-1]]
    return sweep


def tdvp_regions(graph, time_step, *, updater_kwargs, tdvp_order, nsites, **kws):
    plan = []
    for step, weight in enumerate(tdvp_sub_time_steps(tdvp_order), start=1):
        plan += first_order_sweep(graph, weight * time_step, reverse=(step % 2 == 0),
                                  updater_kwargs=updater_kwargs, nsites=nsites, **kws)
    return plan
