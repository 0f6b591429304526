cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r02_run38_pytest_all.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python tools/readme_workload.py 100 > gpurun_out/r02_readme_workload_final.log 2>&1
grep nsites gpurun_out/r02_readme_workload_final.log | cut -c1-200
