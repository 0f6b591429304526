set -x
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_scale.py::test_synthetic_setup_is_bitwise_reproducible -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r02_run23_pytest.log
timeout 200 python tools/readme_workload.py 100 timers > gpurun_out/r02_readme_workload_timers.log 2>&1
tail -2 gpurun_out/r02_readme_workload_timers.log | cut -c1-900
timeout 200 python tools/readme_workload.py 100 > gpurun_out/r02_readme_workload.log 2>&1
tail -2 gpurun_out/r02_readme_workload.log | cut -c1-300
timeout 200 python bench.py --config 1 --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg1_b.json 2> gpurun_out/r02_bench_cfg1_b.err
cut -c1-400 gpurun_out/r02_bench_cfg1_b.json
