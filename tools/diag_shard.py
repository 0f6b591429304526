"""Diagnostic (torchrun, N ranks): sharded vs single-GPU H_eff application vs the naive-FP64-kernel application on the
benchmark problem; prints pairwise max relative differences and how the difference correlates with theta / y."""
import os, sys, json
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
chi, N = int(sys.argv[1]), int(sys.argv[2])
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ctx = ns.Context(local)
ctx.set_option("skip_identity", skip)
net, region = build_problem(chi, N, ctx)
net.extract(region)
th, _ = net.local_download()
# checksum of the replicated inputs across ranks
t = torch.tensor([float(np.abs(th).sum()), float((th * np.arange(th.size).reshape(th.shape, order="F") % 7).sum())], device="cuda", dtype=torch.float64)
t0 = t.clone(); dist.broadcast(t0, src=0)
same_theta = bool(torch.equal(t, t0))
y_rep = net.matvec_device(1, download=True)
ctx.set_option("gemm_impl", 1)
y_naive = net.matvec_device(1, download=True)
ctx.set_option("gemm_impl", 0)
sh = setup_sharded_matvec(net, dist, rank, world)
y_sh = net.matvec_device(1, download=True)
sh.enable(False)
y_rep2 = net.matvec_device(1, download=True)
m = np.abs(y_naive).max()
def rel(a, b): return float(np.abs(a - b).max() / m)
d = y_sh - y_rep
# projections of the difference on theta and on y
c_th = float(np.vdot(th, d) / np.vdot(th, th)); c_y = float(np.vdot(y_rep, d) / np.vdot(y_rep, y_rep))
res_th = float(np.linalg.norm(d - c_th * th) / max(np.linalg.norm(d), 1e-300)); res_y = float(np.linalg.norm(d - c_y * y_rep) / max(np.linalg.norm(d), 1e-300))
import hashlib
def h(a): return hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()[:10]
hs = net.to_host()
hpsi = hashlib.md5(b"".join(np.ascontiguousarray(hs.tensors[v]).tobytes() for v in hs.graph.vertices)).hexdigest()[:10]
print(f"HASH rank={rank} theta={h(th)} psi_all={hpsi} y_rep={h(y_rep)} y_naive={h(y_naive)} y_sh={h(y_sh)} sums y_rep={y_rep.sum()!r} y_sh={y_sh.sum()!r}", flush=True)
print(f"DIAG rank={rank} chi={chi} N={N} skip={skip} same_theta_across_ranks={same_theta} rep_vs_naive={rel(y_rep, y_naive):.2e} "
      f"sh_vs_naive={rel(y_sh, y_naive):.2e} sh_vs_rep={rel(y_sh, y_rep):.2e} rep_vs_rep2={rel(y_rep, y_rep2):.2e} "
      f"d~theta coef {c_th:.2e} resid {res_th:.2f}; d~y coef {c_y:.2e} resid {res_y:.2f}", flush=True)
dist.barrier(); dist.destroy_process_group()
