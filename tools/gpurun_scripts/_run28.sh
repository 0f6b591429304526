cd $GRAFT_REPO_ROOT
for v in 1 0; do
  NSB_DEBUG_EIGH=1 timeout 200 python tools/perf_eigh.py 8192 gauss nocheck eigh_l2_persist=$v > gpurun_out/r02_perf_eigh_l2_$v.log 2>&1
  grep -E "\[eigh\]|factorize_eigh" gpurun_out/r02_perf_eigh_l2_$v.log | cut -c1-260
done
