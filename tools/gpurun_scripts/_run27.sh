cd $GRAFT_REPO_ROOT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/_nccl_shard_worker.py > gpurun_out/r02_worker.log 2>&1
grep -v "^\[W\|^W0\|^\*\*\*\|Setting OMP" gpurun_out/r02_worker.log | grep -B2 -A12 "Traceback" | head -60
tail -3 gpurun_out/r02_worker.log
