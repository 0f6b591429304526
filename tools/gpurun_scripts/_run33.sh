cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -12 | tee gpurun_out/r02_run33_pytest_all.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
NSB_HOST_PROF=1 timeout 200 python tools/readme_workload.py 100 > gpurun_out/r02_readme_hostprof_v2.log 2>&1
grep -E "host-prof|nsites" gpurun_out/r02_readme_hostprof_v2.log | cut -c1-200
timeout 200 python bench.py --config 1 --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg1_c.json 2> gpurun_out/r02_bench_cfg1_c.err
cut -c1-260 gpurun_out/r02_bench_cfg1_c.json
