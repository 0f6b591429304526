set -x
cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -30 | tee gpurun_out/r02_run6_pytest_all.log
timeout 600 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_cfg3.json 2> gpurun_out/r02_bench_cfg3.err
tail -c 1000 gpurun_out/r02_bench_cfg3.err; cut -c1-1500 gpurun_out/r02_bench_cfg3.json
timeout 900 python bench.py --config 1 --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg1.json 2> gpurun_out/r02_bench_cfg1.err
tail -c 1000 gpurun_out/r02_bench_cfg1.err; cut -c1-2500 gpurun_out/r02_bench_cfg1.json
timeout 900 python bench.py --config 4 --steps 10 --warmup 3 > gpurun_out/r02_bench_cfg4.json 2> gpurun_out/r02_bench_cfg4.err
tail -c 1000 gpurun_out/r02_bench_cfg4.err; cut -c1-3500 gpurun_out/r02_bench_cfg4.json
timeout 900 python bench.py --config 5 --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg5.json 2> gpurun_out/r02_bench_cfg5.err
tail -c 1000 gpurun_out/r02_bench_cfg5.err; cut -c1-2500 gpurun_out/r02_bench_cfg5.json
