"""Named tree graphs and traversals used by the region plans (host side).

Stands in for what the reference takes from Graphs.jl / NamedGraphs.jl: `default_root_vertex`,
`post_order_dfs_vertices`, `post_order_dfs_edges`, path queries (src/region_plans/*.jl imports)."""
from __future__ import annotations

from collections import OrderedDict


class NamedGraph:
    """Undirected graph with hashable vertex names; neighbour lists keep edge-insertion order
    (the Euler tour takes the first unvisited neighbour in that order, src/region_plans/euler_tour.jl:5-11)."""

    def __init__(self):
        self._nbrs = OrderedDict()
        self._edges = []

    @property
    def vertices(self):
        return list(self._nbrs.keys())

    @property
    def edges(self):
        return list(self._edges)

    def add_vertex(self, v):
        self._nbrs.setdefault(v, [])
        return self

    def add_edge(self, u, v):
        self.add_vertex(u)
        self.add_vertex(v)
        if v not in self._nbrs[u]:
            self._edges.append((u, v))
            self._nbrs[u].append(v)
            self._nbrs[v].append(u)
        return self

    def neighbors(self, v):
        return list(self._nbrs[v])

    def has_edge(self, u, v):
        return v in self._nbrs[u]

    def leaf_vertices(self):
        return [v for v, n in self._nbrs.items() if len(n) == 1]

    def is_tree(self):
        if len(self._edges) != len(self._nbrs) - 1:
            return False
        verts = self.vertices
        return len(_reachable(self, verts[0])) == len(verts) if verts else True


def _reachable(g, start):
    seen = {start}
    todo = [start]
    while todo:
        x = todo.pop()
        for n in g.neighbors(x):
            if n not in seen:
                seen.add(n)
                todo.append(n)
    return seen


def path_graph(n):
    g = NamedGraph()
    for j in range(1, n + 1):
        g.add_vertex(j)
    for j in range(1, n):
        g.add_edge(j, j + 1)
    return g


def named_comb_tree(tooth_lengths):
    g = NamedGraph()
    for i, ln in enumerate(tooth_lengths, start=1):
        for j in range(1, ln + 1):
            g.add_vertex((i, j))
    for i in range(1, len(tooth_lengths)):
        g.add_edge((i, 1), (i + 1, 1))
    for i, ln in enumerate(tooth_lengths, start=1):
        for j in range(1, ln):
            g.add_edge((i, j), (i, j + 1))
    return g


def star_of_chains(nbranch=3, nbranch_sites=3):
    """The tree of test/utilities/tree_graphs.jl:10-23: centre (0,0) plus `nbranch` chains."""
    g = NamedGraph()
    g.add_vertex((0, 0))
    for b in range(1, nbranch + 1):
        for s in range(1, nbranch_sites + 1):
            g.add_vertex((b, s))
    for b in range(1, nbranch + 1):
        g.add_edge((0, 0), (b, 1))
        for s in range(2, nbranch_sites + 1):
            g.add_edge((b, s - 1), (b, s))
    return g


def default_root_vertex(g):
    return g.leaf_vertices()[-1]


def _dfs(g, root):
    """Returns (post-order vertex list, parent map) of the DFS tree, neighbours in adjacency order."""
    parent = {root: None}
    post = []
    stack = [(root, 0)]
    while stack:
        v, i = stack.pop()
        nb = g.neighbors(v)
        while i < len(nb) and nb[i] in parent:
            i += 1
        if i < len(nb):
            stack.append((v, i + 1))
            parent[nb[i]] = v
            stack.append((nb[i], 0))
        else:
            post.append(v)
    return post, parent


def post_order_dfs_vertices(g, root):
    return _dfs(g, root)[0]


def post_order_dfs_edges(g, root):
    post, parent = _dfs(g, root)
    return [(v, parent[v]) for v in post if parent[v] is not None]


def vertex_path(g, a, b):
    _, parent = _dfs(g, a)
    out = [b]
    while out[-1] != a:
        out.append(parent[out[-1]])
    return out[::-1]
