set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_qn.py -x -q -m gpu 2>&1 | tail -40 | tee gpurun_out/r02_run4_pytest_qn.log
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_kernels.py -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r02_run4_pytest_b.log
