set -x
cd $GRAFT_REPO_ROOT
export NSB_BENCH_NO_ASSERT=1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $T --master-port 29640 bench.py --gpus 2 --steps 3 --warmup 3 --chi 4096 --nsites 100 --no-region-step > gpurun_out/diag2.json 2> gpurun_out/diag2.err
grep SHARD_DIAG gpurun_out/diag2.err
python -c "import json; d=json.loads(open('gpurun_out/diag2.json').read()); print('DIAG2', d['ms_per_step'], d.get('sharded_vs_replicated_max_rel_err'))"
timeout 900 $T --master-port 29641 bench.py --gpus 2 --steps 3 --warmup 3 --chi 4096 --nsites 100 --no-region-step --opt skip_identity=0 > gpurun_out/diag2b.json 2> gpurun_out/diag2b.err
grep SHARD_DIAG gpurun_out/diag2b.err
python -c "import json; d=json.loads(open('gpurun_out/diag2b.json').read()); print('DIAG2b', d['ms_per_step'], d.get('sharded_vs_replicated_max_rel_err'))"
