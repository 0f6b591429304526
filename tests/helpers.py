"""Shared helpers for the parity tests: conversions between the product's host tensors and the
oracle's labelled tensors, and a GPU-availability probe."""
import numpy as np


def cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def to_oracle_graph(g):
    from oracle.graph import NamedGraph
    og = NamedGraph()
    for v in g.vertices:
        og.add_vertex(v)
    for u, v in g.edges:
        og.add_edge(u, v)
    return og


def _olabel(leg, operator=False):
    from oracle.tensor import site, link, oplink
    if leg[0] == "site":
        return site(leg[1], 0)
    if leg[0] == "site_out":
        return site(leg[1], 1)
    return oplink(leg[1], leg[2]) if operator else link(leg[1], leg[2])


def to_oracle_ttn(host, operator=False):
    """networksolvers_b200.HostTTN -> oracle.models.TTN"""
    from oracle.models import TTN
    from oracle.tensor import Tensor
    og = to_oracle_graph(host.graph)
    tensors = {v: Tensor(np.array(host.tensors[v]), [_olabel(l, operator) for l in host.legs[v]]) for v in host.graph.vertices}
    return TTN(og, tensors, ortho_region=list(host.ortho_region) if not operator else [])


def oracle_array(t, legs, operator=False):
    """oracle Tensor -> numpy array with axes ordered like `legs`."""
    return t.array([_olabel(l, operator) for l in legs])


def neel(g, even_up=True):
    out = {}
    for j, v in enumerate(g.vertices, start=1):
        up = (j % 2 == 0) if even_up else (j % 2 == 1)
        out[v] = "Up" if up else "Dn"
    return out


class SweepRecorder:
    """sweep_callback / region_callback that records eigenvalues, truncation errors and bond dimensions."""

    def __init__(self):
        self.energies, self.truncerrs, self.region_energies, self.maxlinkdims = [], [], [], []

    def region(self, problem, **kws):
        self.region_energies.append(getattr(problem, "eigenvalue", None))
        self.truncerrs.append(getattr(problem, "last_truncerr", None))

    def sweep(self, region_iter, **kws):
        p = region_iter.problem
        self.energies.append(getattr(p, "eigenvalue", None))
        self.maxlinkdims.append(p.state.maxlinkdim())


def to_oracle_qn(host):
    """HostTTN.qn (dict) -> oracle.qn.QNInfo"""
    from oracle.qn import QNInfo
    from oracle.tensor import edge_key
    if host.qn is None:
        return None
    link, side = {}, {}
    for (u, v), arr in host.qn["link"].items():
        link[edge_key(u, v)] = np.array(arr)
        side[edge_key(u, v)] = u
    return QNInfo(host.qn["total"], host.qn["site"], link, side)


def to_oracle_ttn_qn(host):
    psi = to_oracle_ttn(host)
    psi.qn = to_oracle_qn(host)
    return psi
