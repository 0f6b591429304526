set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -12 | tee gpurun_out/r02_run21_pytest_all.log
timeout 600 python bench.py --config 3 --steps 10 --warmup 3 > gpurun_out/r02_bench_cfg3.json 2> gpurun_out/r02_bench_cfg3.err
tail -c 600 gpurun_out/r02_bench_cfg3.err; cut -c1-300 gpurun_out/r02_bench_cfg3.json
timeout 100 python -c "import __graft_entry__ as g; g.smoke()"
