cd $GRAFT_REPO_ROOT
for v in 1 2; do
  NSB_DEBUG_EIGH=1 timeout 200 python tools/perf_eigh.py 4096 8192 gauss nocheck eigh_coop_ctas=$v > gpurun_out/r02_perf_eigh_ctas_$v.log 2>&1
  echo "ctas=$v"; grep -E "tridiagonalise" gpurun_out/r02_perf_eigh_ctas_$v.log | cut -c1-200
done
