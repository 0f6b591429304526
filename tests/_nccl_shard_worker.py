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
    if rank == 0:
        print("NCCL_SHARD_OK", world)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
