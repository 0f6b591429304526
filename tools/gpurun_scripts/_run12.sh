cd $GRAFT_REPO_ROOT
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$T --master-port 29650 tools/diag_shard.py 4096 100 0 2>&1 | grep DIAG
$T --master-port 29651 tools/diag_shard.py 2048 100 0 2>&1 | grep DIAG
$T --master-port 29652 tools/diag_shard.py 2048 100 1 2>&1 | grep DIAG
