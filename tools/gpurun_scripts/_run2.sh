set -x
cd $GRAFT_REPO_ROOT
python tools/perf_qr.py > gpurun_out/r02_perf_qr.log 2>&1
cat gpurun_out/r02_perf_qr.log
timeout 1200 python -m pytest tests/test_gpu_sketch.py tests/test_gpu_kernels.py tests/test_gpu_parity_scale.py -x -q -m gpu 2>&1 | tail -40 | tee gpurun_out/r02_run2_pytest_a.log
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -30 | tee gpurun_out/r02_run2_pytest_all.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1_a.json 2> gpurun_out/r02_bench_n1_a.err
tail -c 3000 gpurun_out/r02_bench_n1_a.err
cat gpurun_out/r02_bench_n1_a.json
