"""Worker for test_cpu_multiprocess.py: world_size-2 gloo run of the sharded H_eff partition (host statement).
Each rank contracts its slab of theta along the last bond with the matching rows of the right environment and
the partial results are all-reduced; the result must equal the unsharded oracle matvec on every rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    from networksolvers_b200.parallel import shard_bounds, reference_sharded_matvec
    from oracle.operator_map import optimal_map
    from oracle.projttn import ProjTTN
    from oracle.models import TTN
    from oracle.graph import path_graph
    from oracle.tensor import Tensor, site, link, oplink, contract, noprime
    rng = np.random.default_rng(42)          # same data on every rank (replicated state)
    chi, d, w = int(os.environ.get("NSB_TEST_CHI", "13")), 2, 5
    g = path_graph(4)
    Wt = {v: Tensor(rng.standard_normal((w, w, d, d)), [oplink(v - 1, v), oplink(v, v + 1), site(v, 0), site(v, 1)]) for v in (2, 3)}
    P = ProjTTN(TTN(g, {1: None, 2: Wt[2], 3: Wt[3], 4: None}, ortho_region=[]), pos=[2, 3])
    Lenv = Tensor(rng.standard_normal((chi, w, chi)), [link(1, 2, 0), oplink(1, 2), link(1, 2, 1)])
    Renv = Tensor(rng.standard_normal((chi, w, chi)), [link(3, 4, 0), oplink(3, 4), link(3, 4, 1)])
    P.environments[(1, 2)], P.environments[(4, 3)] = Lenv, Renv
    theta = Tensor(rng.standard_normal((chi, d, d, chi)), [link(1, 2), site(2), site(3), link(3, 4)])
    full = optimal_map(P, theta).array(theta.labels)
    lo, hi = shard_bounds(chi, rank, world)
    th_slab = Tensor(theta.data[..., lo:hi], theta.labels)
    R_slab = Tensor(Renv.data[lo:hi], Renv.labels)
    X = contract(contract(contract(contract(th_slab, Lenv), Wt[2]), Wt[3]), R_slab)
    part = noprime(X).array(theta.labels)

    def allreduce(a):
        t = torch.from_numpy(a)
        dist.all_reduce(t)

    out = reference_sharded_matvec(part, allreduce)
    err = np.abs(out - full).max() / np.abs(full).max()
    assert err < 1e-13, err
    # every rank holds the same (bitwise) result
    t = torch.from_numpy(out.copy())
    ref = t.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(t, ref)
    # AG form (regions swept to the left: the FIRST environment contracts theta's last bond): all-gather the input slabs,
    # split the first contraction along the environment's bra index -- the result is this rank's slab, no reduction.
    # Roles mirrored: Renv is contracted first (over theta's last bond), Lenv last.
    slabs = [torch.zeros(theta.data[..., 0:(hi - lo)].shape, dtype=torch.float64) for _ in range(world)] if chi % world == 0 else None
    if slabs is not None:
        dist.all_gather(slabs, torch.from_numpy(np.ascontiguousarray(theta.data[..., lo:hi])))
        th_full = Tensor(np.concatenate([t.numpy() for t in slabs], axis=-1), theta.labels)
        assert np.array_equal(th_full.data, theta.data)
        R_cols = Tensor(Renv.data[:, :, lo:hi], Renv.labels)
        Y = contract(contract(contract(contract(th_full, R_cols), Wt[3]), Wt[2]), Lenv)
        out_slab = noprime(Y).array(theta.labels)
        err2 = np.abs(out_slab - full[..., lo:hi]).max() / np.abs(full).max()
        assert err2 < 1e-13, err2
    if rank == 0:
        print("GLOO_SHARD_OK", world, err)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
