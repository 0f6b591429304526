set -x
cd $GRAFT_REPO_ROOT
NSB_HOST_PROF=1 timeout 200 python tools/readme_workload.py 100 > gpurun_out/r02_readme_hostprof.log 2>&1
grep -E "host-prof|nsites" gpurun_out/r02_readme_hostprof.log | cut -c1-220
