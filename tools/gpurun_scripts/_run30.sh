cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r02_run30_pytest.log
timeout 400 ncu --set full --import-source on --clock-control none -k regex:trd_panel_sym --launch-skip 20 --launch-count 1 -f -o gpurun_out/r02_trd_sym_n4096 python tools/perf_eigh.py 4096 gauss nocheck > gpurun_out/r02_ncu_trd.log 2>&1
tail -3 gpurun_out/r02_ncu_trd.log
ls -la gpurun_out/*.ncu-rep
