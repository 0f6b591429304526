"""Projected tree-tensor-network operator (oracle; test-only).

UPSTREAM `itn.ProjTTN` / `position` / `environment` / `incident_edges`,
used by the reference at src/operator_map.jl:4-5,17-37, src/extracter.jl:14, src/applyexp.jl:38,
src/subspace/densitymatrix.jl:19,39-46.  An environment is keyed by the directed edge (u, v):
the contraction of everything on u's side, pointing into v.

Cache validity: every environment remembers the state-tensor objects it was built from and is
dropped by `position` when any of them has been replaced.  On every sweep plan the reference
generates this coincides with the upstream invalidation rule (internal edges / reverse of incident
edges), and it is also safe for arbitrary region jumps."""
from __future__ import annotations

from .tensor import contract, dag, prime


class ProjTTN:
    def __init__(self, operator, pos=None, environments=None, deps=None):
        self.operator = operator                  # TTN of W tensors
        self.pos = list(pos) if pos is not None else []   # list of vertices, or [("edge", u, v)]
        self.environments = dict(environments or {})
        self.deps = dict(deps or {})

    @property
    def graph(self):
        return self.operator.graph

    def on_edge(self):
        p = self.pos
        return len(p) == 1 and isinstance(p[0], tuple) and len(p[0]) == 3 and p[0][0] == "edge"

    def sites(self):
        return [] if self.on_edge() else list(self.pos)

    def incident_edges(self):
        g = self.graph
        if self.on_edge():
            _, u, v = self.pos[0]
            return [(u, v), (v, u)]
        out = []
        for v in self.pos:
            for n in g.neighbors(v):
                if n not in self.pos:
                    out.append((n, v))
        return out

    def environment(self, e):
        return self.environments[e]

    def copy(self):
        return ProjTTN(self.operator, self.pos, self.environments, self.deps)


def make_environment(P: ProjTTN, psi, e, counter=None):
    """env(u->v) = psi[u] * W[u] * dag(prime(psi[u])) * prod env(n->u), n != v  (App. A.2)."""
    if e in P.environments:
        return P
    u, v = e
    g = P.graph
    others = [n for n in g.neighbors(u) if n != v]
    for n in others:
        P = make_environment(P, psi, (n, u), counter)
    A = psi[u]
    T = A
    envs = [P.environments[(n, u)] for n in others]
    # upstream heuristic order: two environments, operator, bra, remaining environments
    for env in envs[:2]:
        T = contract(T, env)
    T = contract(T, P.operator[u])
    T = contract(T, dag(prime(A)))
    for env in envs[2:]:
        T = contract(T, env)
    P.environments[e] = T
    dep = {u: A}
    for n in others:
        dep.update(P.deps[(n, u)])
    P.deps[e] = dep
    if counter is not None:
        counter["env_builds"] = counter.get("env_builds", 0) + 1
    return P


def position(P: ProjTTN, psi, region, counter=None):
    """UPSTREAM `itn.position(P, psi, region)`; region = list of vertices or [("edge", u, v)]."""
    Q = P.copy()
    Q.pos = list(region)
    keep = {}
    for e, env in Q.environments.items():
        if all(psi[x] is t for x, t in Q.deps[e].items()):
            keep[e] = env
    Q.environments = keep
    Q.deps = {e: Q.deps[e] for e in keep}
    for e in Q.incident_edges():
        Q = make_environment(Q, psi, e, counter)
    return Q
