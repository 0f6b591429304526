cd $GRAFT_REPO_ROOT
timeout 200 python tools/perf_small_svd2.py 2>&1 | tee gpurun_out/r02_perf_dsmem_spc.log | cut -c1-300
