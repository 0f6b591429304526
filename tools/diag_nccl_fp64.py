"""Diagnostic: accuracy of NCCL FP64 collectives on this box (reduce_scatter / all_reduce of FP64 sums against the exact sum
computed from all-gathered inputs).  torchrun --nproc-per-node N tools/diag_nccl_fp64.py"""
import os
import torch
import torch.distributed as dist

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
r, w = dist.get_rank(), dist.get_world_size()
g = torch.Generator(device="cuda").manual_seed(100 + r)
n = 1 << 24
x = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
allx = [torch.empty_like(x) for _ in range(w)]
dist.all_gather(allx, x)
exact = torch.stack(allx).sum(0)
y = x.clone()
dist.all_reduce(y)
e_ar = ((y - exact).abs().max() / exact.abs().max()).item()
out = torch.empty(n // w, dtype=torch.float64, device="cuda")
dist.reduce_scatter_tensor(out, x)
e_rs = ((out - exact[r * (n // w):(r + 1) * (n // w)]).abs().max() / exact.abs().max()).item()
if r == 0:
    print(f"NCCL_FP64 world={w} NVLS={os.environ.get('NCCL_NVLS_ENABLE', 'default')} ALGO={os.environ.get('NCCL_ALGO', 'default')} "
          f"allreduce_err={e_ar:.3e} reduce_scatter_err={e_rs:.3e}", flush=True)
dist.barrier()
dist.destroy_process_group()
