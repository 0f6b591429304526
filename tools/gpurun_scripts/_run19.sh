cd $GRAFT_REPO_ROOT
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$T --master-port 29680 tools/diag_shard2.py 2048 100 12 2>&1 | grep -E "DIAG2" | cut -c1-700
