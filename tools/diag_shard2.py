"""Diagnostic (torchrun): after ONE set-up, apply the sharded H_eff many times and compare each result with the single-GPU
application; then the same with host synchronisation around every collective."""
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
net, region = build_problem(chi, N, ctx)
net.extract(region)
y_rep = net.matvec_device(1, download=True)
m = np.abs(y_rep).max()
sh = setup_sharded_matvec(net, dist, rank, world)
for mode in (0, 1):
    ctx.set_option("nccl_sync", mode)
    errs, hashes = [], []
    for i in range(reps):
        y = net.matvec_device(1, download=True)
        errs.append(float(np.abs(y - y_rep).max() / m))
        hashes.append(hashlib.md5(y.tobytes()).hexdigest()[:6])
    print(f"DIAG2 rank={rank} nccl_sync={mode} errs={['%.1e' % e for e in errs]} hashes={hashes}", flush=True)
dist.barrier(); dist.destroy_process_group()
