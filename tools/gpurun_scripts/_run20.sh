cd $GRAFT_REPO_ROOT
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 500 $T --master-port 29690 tools/diag_shard3.py 2048 100 8 2>&1 | grep -E "DIAG3|Error|error" | cut -c1-600
