cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r02_run36_pytest_all.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python tools/readme_workload.py 100 > gpurun_out/r02_readme_workload_final.log 2>&1
grep nsites gpurun_out/r02_readme_workload_final.log | cut -c1-200
timeout 300 python bench.py --no-full-sweep --steps 10 --warmup 3 > gpurun_out/r02_bench_nofullsweep.json 2> gpurun_out/r02_bench_nofullsweep.err
tail -c 300 gpurun_out/r02_bench_nofullsweep.err; cut -c1-200 gpurun_out/r02_bench_nofullsweep.json
