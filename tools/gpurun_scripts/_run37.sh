cd $GRAFT_REPO_ROOT
timeout 120 python tools/perf_small_gemm.py 2>&1 | tee gpurun_out/r02_perf_small_gemm.log | cut -c1-220
