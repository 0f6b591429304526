cd $GRAFT_REPO_ROOT
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for i in 1 2 3 4; do
$T --master-port 2967$i tools/diag_shard.py 2048 100 1 2>&1 | grep -E "DIAG" | cut -c1-200
done
