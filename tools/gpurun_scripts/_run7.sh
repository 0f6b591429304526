set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29610 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
tail -c 1500 gpurun_out/r02_bench_n8.err
cat gpurun_out/r02_bench_n8.json
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_hotpath.py -q -m gpu -k "multi or sharded" 2>&1 | tail -8
