set -x
cd $GRAFT_REPO_ROOT
export NSB_BENCH_NO_ASSERT=1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
B="bench.py --gpus 8 --steps 3 --warmup 3 --chi 4096 --nsites 26 --no-region-step"
for o in "" "--opt skip_identity_sharded=0" "--opt skip_identity=0"; do
  timeout 600 $T --master-port 29630 $B $o > gpurun_out/diag8.json 2> gpurun_out/diag8.err
  python -c "import json; d=json.loads(open('gpurun_out/diag8.json').read()); print('DIAG8', '$o', d['ms_per_step'], d.get('sharded_vs_replicated_max_rel_err'))"
done
T4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 600 $T4 --master-port 29631 bench.py --gpus 4 --steps 3 --warmup 3 --chi 4096 --nsites 26 --no-region-step > gpurun_out/diag4.json 2> gpurun_out/diag4.err
python -c "import json; d=json.loads(open('gpurun_out/diag4.json').read()); print('DIAG4', d['ms_per_step'], d.get('sharded_vs_replicated_max_rel_err'))"
