cd $GRAFT_REPO_ROOT
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$T --master-port 29660 tools/diag_shard.py 2048 100 1 2>&1 | grep -E "DIAG|HASH"
$T --master-port 29661 tools/diag_shard.py 2048 100 1 2>&1 | grep -E "DIAG|HASH"
