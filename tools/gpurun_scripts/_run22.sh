set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
grep -E "sharded-vs-single|Error" gpurun_out/r02_bench_n8.err | head -10
cut -c1-300 gpurun_out/r02_bench_n8.json
