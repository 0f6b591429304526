cd $GRAFT_REPO_ROOT
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_eigh.py -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/r02_run32_pytest.log
NSB_DEBUG_EIGH=1 timeout 200 python tools/perf_eigh.py 4096 8192 gauss nocheck > gpurun_out/r02_perf_eigh_v2.log 2>&1
grep -E "tridiagonalise" gpurun_out/r02_perf_eigh_v2.log | cut -c1-200
timeout 200 python tools/perf_small_svd.py 2>&1 | tee gpurun_out/r02_perf_small_svd_v2.log | cut -c1-300
