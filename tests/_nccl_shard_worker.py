"""Worker for the multi-GPU test: sharded H_eff application (theta split along its last bond + NCCL all-reduce)
must reproduce the single-GPU matvec and Ritz value on every rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    import networksolvers_b200 as ns
    from networksolvers_b200.parallel import setup_sharded_matvec
    ctx = ns.Context(local)
    comm_ready = False
    for cplx, fused in ((False, False), (True, False), (False, True), (True, True)):
        g = ns.path_graph(12)
        sites = ns.siteinds("S=1/2", g)
        H = ns.ttno(ns.heisenberg(g), sites)
        psi = ns.random_state(sites, 48, seed=9, dtype=complex if cplx else float)
        net = ns.EigsolveProblem(state=psi, operator=H, ctx=ctx).net
        net.extract([6, 7])
        ref = net.matvec_device(1, download=True)
        val_ref, _ = net.update_eigsolve()
        # same problem again, sharded
        net2 = ns.EigsolveProblem(state=psi, operator=H, ctx=ctx).net
        net2.extract([6, 7])
        sh = setup_sharded_matvec(net2, dist, rank, world, fused=fused, init=not comm_ready)
        comm_ready = True
        assert sh.active
        out = net2.matvec_device(1, download=True)
        err = np.abs(out - ref).max() / np.abs(ref).max()
        assert err < 1e-13, err
        val, info = net2.update_eigsolve()
        assert info.nmatvec == 3 and abs(val - val_ref) < 1e-12 * max(1, abs(val_ref)), (val, val_ref)
        ins = net2.insert((1e-10, 1, 40))
        t = torch.tensor([val, float(ins.newdim), ins.truncerr], dtype=torch.float64, device="cuda")
        t0 = t.clone()
        dist.broadcast(t0, src=0)
        assert torch.equal(t, t0), "ranks diverged"
    # ---- whole region steps in both sweep directions, every phase sharded (Krylov vectors as slabs with reduce-scatter /
    # all-gather applications, environment update split + all-reduce, factorisation by column slabs + all-gather) against
    # the same steps on one GPU.  eigh_min_n is lowered so that the Gram + eigh route (the one that is split) runs at this size.
    ctx.set_option("eigh_min_n", 128)
    try:
        for cplx in (False, True):
            g = ns.path_graph(16)
            sites = ns.siteinds("S=1/2", g)
            H = ns.ttno(ns.heisenberg(g), sites)
            psi = ns.random_state(sites, 128, seed=5, dtype=complex if cplx else float)
            nets = [ns.EigsolveProblem(state=psi, operator=H, ctx=ctx).net for _ in range(2)]
            for n_ in nets:
                n_.extract([8, 9])
            sh = setup_sharded_matvec(nets[1], dist, rank, world, fused=False, init=False)
            assert sh.active
            regions = [[8, 9], [9, 10], [10, 11], [11, 10], [10, 9], [9, 8]]
            for reg in regions:
                res = []
                for n_ in nets:
                    n_.extract(reg)
                    th, _ = n_.local_download()
                    hv = n_.matvec_host(th)
                    # distributed host-buffer call: this rank's slab in, the matching slab of theta' out
                    lo, hi, d = n_.shard_range()
                    assert (n_ is nets[0]) == ((lo, hi) == (0, d)) or world == 1, (lo, hi, d)
                    hs = n_.matvec_host_slab(th[..., lo:hi])
                    assert np.abs(hs - hv[..., lo:hi]).max() <= 1e-12 * np.abs(hv).max(), (reg, "slab call")
                    val, info = n_.update_eigsolve()
                    th2, _ = n_.local_download()
                    ins = n_.insert((1e-12, 1, 128))
                    res.append((th, hv, val, th2, ins.newdim, ins.truncerr, info.nmatvec))
                a, b = res
                # gauge-invariant comparisons (the two factorisation routes may differ by signs of basis vectors)
                ea, eb = np.vdot(a[0], a[1]).real / np.vdot(a[0], a[0]).real, np.vdot(b[0], b[1]).real / np.vdot(b[0], b[0]).real
                assert abs(ea - eb) <= 1e-11 * max(1.0, abs(ea)), (reg, "energy expectation", ea, eb)
                assert abs(np.linalg.norm(a[0]) - np.linalg.norm(b[0])) <= 1e-11 * np.linalg.norm(a[0]), (reg, "theta norm")
                assert abs(a[2] - b[2]) <= 1e-11 * max(1.0, abs(a[2])), (reg, a[2], b[2])
                assert abs(np.linalg.norm(a[3]) - 1.0) <= 1e-12 and abs(np.linalg.norm(b[3]) - 1.0) <= 1e-12
                assert a[4] == b[4] and abs(a[5] - b[5]) <= 1e-10 and a[6] == b[6] == 3, (reg, a[4:], b[4:])
            t = torch.tensor([res[1][2], res[1][5]], dtype=torch.float64, device="cuda")
            t0 = t.clone()
            dist.broadcast(t0, src=0)
            assert torch.equal(t, t0), "ranks diverged"
            # environments away from the region live as 1 / G slabs on the sharded network, in full on the other
            r0, f0 = nets[0].env_bytes()
            r1, f1 = nets[1].env_bytes()
            assert r0 == f0 and f1 == f0 and r1 < f1, (r0, f0, r1, f1)
            # ... and come back bit for bit: the energy expectation at the far end of the chain agrees with the replica
            for reg in ([3, 2], [2, 3]):
                vals = []
                for n_ in nets:
                    n_.extract(reg)
                    th, _ = n_.local_download()
                    hv = n_.matvec_host(th)
                    vals.append(np.vdot(th, hv).real / np.vdot(th, th).real)
                assert abs(vals[0] - vals[1]) <= 1e-11 * max(1.0, abs(vals[0])), (reg, vals)
    finally:
        ctx.set_option("eigh_min_n", 1024)
    if rank == 0:
        print("NCCL_SHARD_OK", world)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
