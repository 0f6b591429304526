"""Variational fitting of a tree tensor network (oracle; test-only): restatement of src/fitting.jl.

The reference fits a ket |psi> to a target |x> (`truncate`, src/fitting.jl:90-97) or A|x> (`apply`, :99-112) by
sweeping over the ket: the local tensor of a region is the environment of the overlap network <psi| A |x> with the
region's ket tensors removed (src/fitting.jl:25-40; on a tree the belief-propagation cache it uses is the exact
contraction), the updater only records the overlap (:42-49), the common inserter writes the tensor back with
`normalize` and `set_orthogonal_region=false` (:78).

Environments are keyed by directed edges like the projected operator's (projttn.py): env(u->v) = x[u] * A[u] *
dag(prime(psi[u])) * prod env(n->u); the target's links carry prime level 2 so that they never collide with the
ket's.  With dag() taken on the ket layer the contracted region environment is directly the new ket tensor (the
reference conjugates at the end instead, src/fitting.jl:38)."""
from __future__ import annotations

import numpy as np

from .gauge import orthogonalize
from .graph import NamedGraph
from .models import TTN
from .region_plans import euler_sweep
from .tensor import Tensor, contract, dag, link, oplink, prime, site

XPLEV = 2


def _xlabel(l):
    return ("l", l[1], XPLEV) if l[0] == "l" else l


def identity_ttno(g: NamedGraph, d, dtype=float):
    """Operator network with all links of dimension 1 acting as the identity (fit to |x> itself)."""
    tensors = {}
    for v in g.vertices:
        labels = [site(v, 0), site(v, 1)] + [oplink(v, n) for n in g.neighbors(v)]
        tensors[v] = Tensor(np.eye(d, dtype=dtype).reshape([d, d] + [1] * len(g.neighbors(v))), labels)
    return TTN(g, tensors, ortho_region=[])


def delta_state(g: NamedGraph, d, link_space, dtype=float):
    """`ITensorNetwork(v -> inds -> delta(inds), siteinds; link_space)` (src/fitting.jl:91-93): every tensor is 1 where
    all of its indices agree."""
    tensors = {}
    for v in g.vertices:
        nbrs = g.neighbors(v)
        labels = ([link(v, nbrs[0])] if nbrs else []) + [site(v)] + [link(v, n) for n in nbrs[1:]]
        shape = [d if l[0] == "s" else link_space for l in labels]
        data = np.zeros(shape, dtype=dtype)
        for i in range(min(shape)):
            data[(i,) * len(shape)] = 1.0
        tensors[v] = Tensor(data, labels)
    return TTN(g, tensors)


def random_tensornetwork(g: NamedGraph, d, link_space, rng, dtype=float):
    """`itn.random_tensornetwork(rng, elt, s; link_space)`: i.i.d. normal entries, uniform link dimension."""
    tensors = {}
    for v in g.vertices:
        nbrs = g.neighbors(v)
        labels = ([link(v, nbrs[0])] if nbrs else []) + [site(v)] + [link(v, n) for n in nbrs[1:]]
        shape = [d if l[0] == "s" else link_space for l in labels]
        data = rng.standard_normal(shape)
        if np.issubdtype(np.dtype(dtype), np.complexfloating):
            data = data + 1j * rng.standard_normal(shape)
        tensors[v] = Tensor(data.astype(dtype), labels)
    return TTN(g, tensors)


class FittingProblem:
    """src/fitting.jl:8-18.  `state` is the ket being fitted; the overlap network is (state, operator, target)."""

    def __init__(self, state, target, operator, overlap=0.0, envs=None, deps=None):
        self.state, self.target, self.operator, self.overlap = state, target, operator, overlap
        self.envs = dict(envs or {})
        self.deps = dict(deps or {})

    def setproperties(self, **kw):
        new = FittingProblem(self.state, self.target, self.operator, self.overlap, self.envs, self.deps)
        for k, v in kw.items():
            setattr(new, k, v)
        return new


def _make_env(F: FittingProblem, psi, e):
    if e in F.envs:
        return
    u, v = e
    others = [n for n in psi.graph.neighbors(u) if n != v]
    for n in others:
        _make_env(F, psi, (n, u))
    T = F.target[u]
    for n in others:
        T = contract(T, F.envs[(n, u)])
    T = contract(T, F.operator[u])
    T = contract(T, dag(prime(psi[u])))
    F.envs[e] = T
    dep = {u: psi[u]}
    for n in others:
        dep.update(F.deps[(n, u)])
    F.deps[e] = dep


def extracter(problem: FittingProblem, region_iter, *, sweep, **kws):
    """src/fitting.jl:25-40: gauge walk to the region, environment update along the path, region environment."""
    region = region_iter.current_region()
    psi = orthogonalize(problem.state, region)
    F = problem.setproperties(state=psi)
    keep = {e: env for e, env in F.envs.items() if all(psi[x] is t for x, t in F.deps[e].items())}
    F.envs = keep
    F.deps = {e: F.deps[e] for e in keep}
    g = psi.graph
    local = None
    for v in region:
        for n in g.neighbors(v):
            if n not in region:
                _make_env(F, psi, (n, v))
    for v in region:
        t = contract(F.target[v], F.operator[v])
        local = t if local is None else contract(local, t)
        for n in g.neighbors(v):
            if n not in region:
                local = contract(local, F.envs[(n, v)])
    local = Tensor(local.data, tuple((l[0], l[1], 0) for l in local.labels))   # noprime
    return F, local


def updater(F: FittingProblem, local_tensor, region_iter, *, outputlevel, **kws):
    """src/fitting.jl:42-49."""
    n = float(np.real(np.vdot(local_tensor.data, local_tensor.data)))
    F = F.setproperties(overlap=n / np.sqrt(n) if n > 0 else 0.0)
    if outputlevel >= 2:
        print("  Region %s: squared overlap = %.12f" % (region_iter.current_region(), F.overlap))
    return F, local_tensor


def region_plan(F: FittingProblem, *, nsites, **sweep_kwargs):
    """src/fitting.jl:51-53."""
    return euler_sweep(F.state.graph, nsites=nsites, **sweep_kwargs)


def fit_tensornetwork(target, operator, init_state, *, nsweeps=25, nsites=1, outputlevel=0, extracter_kwargs=None,
                      updater_kwargs=None, inserter_kwargs=None, normalize=True, **kws):
    """src/fitting.jl:55-84.  `target` |x>, `operator` A (or None for the identity), `init_state` the initial ket."""
    from .sweep import sweep_iterator, sweep_solve
    g = init_state.graph
    d = init_state[g.vertices[0]].dim(site(g.vertices[0]))
    dtype = np.result_type(*[target[v].data.dtype for v in g.vertices], *([operator[v].data.dtype for v in g.vertices] if operator else []))
    if operator is None:
        operator = identity_ttno(g, d, dtype)
    x = TTN(g, {v: target[v].relabel({l: _xlabel(l) for l in target[v].labels}) for v in g.vertices}, ortho_region=[])
    init = TTN(g, {v: Tensor(init_state[v].data.astype(dtype), init_state[v].labels) for v in g.vertices},
               ortho_region=list(g.vertices))
    prob = FittingProblem(init, x, operator)
    ik = dict(inserter_kwargs or {})
    ik.update(normalize=normalize, set_orthogonal_region=False)
    common = dict(nsites=nsites, outputlevel=outputlevel, extracter_kwargs=dict(extracter_kwargs or {}),
                  updater_kwargs=dict(updater_kwargs or {}), inserter_kwargs=ik)
    it = sweep_iterator(prob, [dict(common) for _ in range(nsweeps)])   # the iterator supplies sweep = 1, 2, ...
    conv = sweep_solve(it, outputlevel=outputlevel, **kws)
    return conv.state


def truncate(tn, *, maxdim, cutoff=0.0, **kws):
    """`itn.truncate(tn; maxdim, cutoff)` (src/fitting.jl:90-97)."""
    g = tn.graph
    d = tn[g.vertices[0]].dim(site(g.vertices[0]))
    init = delta_state(g, d, maxdim, tn[g.vertices[0]].data.dtype)
    return fit_tensornetwork(tn, None, init, inserter_kwargs=dict(trunc=dict(cutoff=cutoff, maxdim=maxdim)), **kws)


def apply(A, x, *, maxdim, cutoff=0.0, **kws):
    """`itn.apply(A, x; maxdim, cutoff)` (src/fitting.jl:99-112)."""
    g = x.graph
    d = x[g.vertices[0]].dim(site(g.vertices[0]))
    init = delta_state(g, d, maxdim, x[g.vertices[0]].data.dtype)
    return fit_tensornetwork(x, A, init, inserter_kwargs=dict(trunc=dict(cutoff=cutoff, maxdim=maxdim)), **kws)


def inner(a, b, A=None):
    """Exact <a| A |b> by dense contraction (`itn.inner(...; alg="exact")`)."""
    from .ed import state_vector
    va, vb = state_vector(a), state_vector(b)
    if A is None:
        return np.vdot(va, vb)
    from .ed import ttno_dense
    g = a.graph
    d = a[g.vertices[0]].dim(site(g.vertices[0]))
    return np.vdot(va, ttno_dense(A, g, d) @ vb)
