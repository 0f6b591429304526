"""Diagnostic (torchrun): several complete builds of the benchmark problem in every rank; per build the hashes of every site
tensor, of the single-GPU application and of the sharded application are compared ACROSS ranks."""
import os, sys, hashlib
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import networksolvers_b200 as ns
from networksolvers_b200.parallel import setup_sharded_matvec
from bench import build_problem
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
chi, N, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ctx = ns.Context(local)
h = lambda a: hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()[:8]
first = True
for rep in range(reps):
    net, region = build_problem(chi, N, ctx)
    net.extract(region)
    y_rep = net.matvec_device(1, download=True)
    hs = net.to_host()
    site_h = [h(hs.tensors[v]) for v in hs.graph.vertices]
    sh = setup_sharded_matvec(net, dist, rank, world, init=first)
    first = False
    y_sh = net.matvec_device(1, download=True)
    err = float(np.abs(y_sh - y_rep).max() / np.abs(y_rep).max())
    rec = dict(rank=rank, y_rep=h(y_rep), y_sh=h(y_sh), sites=site_h, err=err)
    allrec = [None] * world
    dist.all_gather_object(allrec, rec)
    if rank == 0:
        r0 = allrec[0]
        msg = []
        for r in allrec[1:]:
            bad = [i for i, (a, b) in enumerate(zip(r0["sites"], r["sites"])) if a != b]
            msg.append(f"rank{r['rank']}: y_rep {'=' if r['y_rep'] == r0['y_rep'] else 'DIFF'} y_sh {'=' if r['y_sh'] == r0['y_sh'] else 'DIFF'} differing sites {bad[:10]} (of {len(bad)})")
        print(f"DIAG3 build {rep}: errs {[('%.1e' % r['err']) for r in allrec]} y_rep0 {r0['y_rep']} y_sh0 {r0['y_sh']}; " + "; ".join(msg), flush=True)
    net.close()
dist.barrier(); dist.destroy_process_group()
