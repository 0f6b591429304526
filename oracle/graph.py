"""Minimal named-graph helpers for the oracle (test-only).

Restates the graph utilities the reference takes from Graphs.jl / NamedGraphs.jl (UPSTREAM,
SURVEY.md App. A.10) and the graphs its tests build (test/utilities/tree_graphs.jl:10-23,
test/tdvp/test_tree_tdvp.jl:10-22).
"""
from __future__ import annotations


class NamedGraph:
    def __init__(self):
        self.vertices = []          # insertion order
        self.edges = []             # (src, dst) insertion order
        self._adj = {}

    def add_vertex(self, v):
        if v not in self._adj:
            self.vertices.append(v)
            self._adj[v] = []

    def add_edge(self, u, v):
        self.add_vertex(u)
        self.add_vertex(v)
        if v in self._adj[u]:
            return
        self.edges.append((u, v))
        self._adj[u].append(v)
        self._adj[v].append(u)

    def neighbors(self, v):
        return list(self._adj[v])

    def degree(self, v):
        return len(self._adj[v])

    def leaf_vertices(self):
        return [v for v in self.vertices if self.degree(v) == 1]

    def has_edge(self, u, v):
        return v in self._adj[u]


def path_graph(n):
    g = NamedGraph()
    for j in range(1, n + 1):
        g.add_vertex(j)
    for j in range(1, n):
        g.add_edge(j, j + 1)
    return g


def build_tree(nbranch=3, nbranch_sites=3):
    """test/utilities/tree_graphs.jl:10-23 -- centre (0,0) with nbranch chains."""
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


def chain_plus_ancilla(nchain):
    """test/tdvp/test_tree_tdvp.jl:10-22."""
    g = NamedGraph()
    for j in range(1, nchain + 1):
        g.add_vertex(j)
    for j in range(1, nchain):
        g.add_edge(j, j + 1)
    g.add_vertex(0)
    g.add_edge(0, nchain // 2)
    return g


def named_comb_tree(tooth_lengths):
    """UPSTREAM NamedGraphGenerators.named_comb_tree: backbone (i,1), tooth i = (i,1)..(i,len_i)."""
    g = NamedGraph()
    nx = len(tooth_lengths)
    for i in range(1, nx + 1):
        for j in range(1, tooth_lengths[i - 1] + 1):
            g.add_vertex((i, j))
    for i in range(1, nx):
        g.add_edge((i, 1), (i + 1, 1))
    for i in range(1, nx + 1):
        for j in range(1, tooth_lengths[i - 1]):
            g.add_edge((i, j), (i, j + 1))
    return g


def default_root_vertex(g):
    """UPSTREAM (believed) `last(leaf_vertices(g))` (README.md:35)."""
    return g.leaf_vertices()[-1]


def dfs_parents(g, root):
    """DFS tree from root, neighbours in adjacency order.  Returns (preorder, parent dict)."""
    parent = {root: None}
    order = []
    stack = [(root, iter(g.neighbors(root)))]
    order.append(root)
    while stack:
        v, it = stack[-1]
        for n in it:
            if n not in parent:
                parent[n] = v
                order.append(n)
                stack.append((n, iter(g.neighbors(n))))
                break
        else:
            stack.pop()
    return order, parent


def post_order_dfs_vertices(g, root):
    out = []
    parent = {root: None}

    def rec(v):
        for n in g.neighbors(v):
            if n not in parent:
                parent[n] = v
                rec(n)
        out.append(v)

    import sys

    sys.setrecursionlimit(max(10000, sys.getrecursionlimit()))
    rec(root)
    return out


def post_order_dfs_edges(g, root):
    """Edges child->parent in DFS post-order (regions move toward the root)."""
    _, parent = dfs_parents(g, root)
    return [(v, parent[v]) for v in post_order_dfs_vertices(g, root) if parent[v] is not None]


def tree_path(g, a, b):
    """Vertex path a..b in the tree."""
    _, parent = dfs_parents(g, a)
    path = [b]
    while path[-1] != a:
        path.append(parent[path[-1]])
    return path[::-1]


def subtree_side(g, u, v):
    """Vertices on u's side of the edge (u, v)."""
    seen = {u}
    stack = [u]
    while stack:
        x = stack.pop()
        for n in g.neighbors(x):
            if n not in seen and not (x == u and n == v):
                seen.add(n)
                stack.append(n)
    return seen
