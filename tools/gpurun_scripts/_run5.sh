set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_qn.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r02_run5_pytest_qn.log
timeout 600 python bench.py --config 3 --chi 512 --nsites 24 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_cfg3_chi512.json 2> gpurun_out/r02_bench_cfg3_chi512.err
tail -c 2000 gpurun_out/r02_bench_cfg3_chi512.err; cat gpurun_out/r02_bench_cfg3_chi512.json
timeout 1500 python bench.py --config 3 --steps 10 --warmup 3 > gpurun_out/r02_bench_cfg3.json 2> gpurun_out/r02_bench_cfg3.err
tail -c 2000 gpurun_out/r02_bench_cfg3.err; cat gpurun_out/r02_bench_cfg3.json
