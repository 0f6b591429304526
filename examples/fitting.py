"""examples/fitting.jl of the reference through networksolvers_b200 (needs a B200): `truncate` and `apply` on a comb tree by
fitting sweeps, with the reference's fidelity assertions (FP64 / complex128: the device library has no single precision).

    python examples/fitting.py
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import networksolvers_b200 as ns  # noqa: E402
from quench_evolution import dense_hamiltonian  # noqa: E402


def fidelity(a, b):
    va, vb = a.to_dense(), b.to_dense()
    return np.vdot(va, vb) / np.sqrt(np.vdot(va, va) * np.vdot(vb, vb))


def fitting(dry_run=False):
    for elt in (np.float64, np.complex128):
        print()
        print("elt =", elt.__name__)
        eps = np.finfo(np.float64).eps
        g = ns.named_comb_tree((3, 2))
        s = ns.siteinds("S=1/2", g)
        rng = np.random.default_rng(1234)
        a = ns.random_tensornetwork(s, 3, rng, elt)
        H = ns.ttno(ns.heisenberg(g), s, dtype=elt)
        if dry_run:
            print(f"fitting: {len(g.vertices)} vertices, maxlinkdim(a) = {a.maxlinkdim()}, operator link dimension {H.maxlinkdim()}")
            continue
        # one-site truncation
        b = ns.truncate(a, maxdim=3).to_host()
        f = fidelity(a, b)
        print("One-site truncation. Fidelity =", f)
        assert abs(abs(f) - 1.0) <= 50 * eps
        # two-site truncation
        a = ns.random_tensornetwork(s, 3, rng, elt)
        b = ns.truncate(a, maxdim=3, cutoff=1e-16, nsites=2).to_host()
        f = fidelity(a, b)
        print("Two-site truncation. Fidelity =", f)
        assert abs(abs(f) - 1.0) <= 50 * eps
        # one-site / two-site apply (no normalisation)
        Hd = dense_hamiltonian(ns.heisenberg(g), s)
        for nsites, kw in ((1, {}), (2, dict(cutoff=1e-16))):
            a = ns.random_tensornetwork(s, 2, rng, elt)
            Ha = ns.apply(H, a, maxdim=4, nsites=nsites, normalize=False, **kw).to_host()
            va = a.to_dense()
            f = np.vdot(Ha.to_dense(), va) / np.vdot(va, Hd @ va)
            print(f"{'One' if nsites == 1 else 'Two'}-site apply. Fidelity =", f)
            assert abs(f - 1.0) <= 200 * eps


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--dry-run", action="store_true")
    fitting(ap.parse_args().dry_run)
