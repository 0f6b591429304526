set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
python tools/perf_qr.py > gpurun_out/r02_perf_qr.log 2>&1
cat gpurun_out/r02_perf_qr.log
timeout 900 python -m pytest tests/test_gpu_sketch.py tests/test_gpu_kernels.py tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -40 | tee gpurun_out/r02_run3_pytest_a.log
timeout 900 python -m pytest tests/test_gpu_hotpath.py -q -m gpu -k "sharded" 2>&1 | tail -40 | tee gpurun_out/r02_run3_pytest_nccl.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus 2 --steps 10 --warmup 3 --chi 2048 --nsites 40 > gpurun_out/r02_bench_n2_chi2048.json 2> gpurun_out/r02_bench_n2_chi2048.err
tail -c 2500 gpurun_out/r02_bench_n2_chi2048.err
cat gpurun_out/r02_bench_n2_chi2048.json
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --chi 2048 --nsites 40 --no-full-sweep --no-cpu-baseline > gpurun_out/r02_bench_n1_chi2048.json 2> gpurun_out/r02_bench_n1_chi2048.err
cat gpurun_out/r02_bench_n1_chi2048.json
