set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
python tools/perf_qr.py > gpurun_out/r02_perf_qr.log 2>&1
cat gpurun_out/r02_perf_qr.log
timeout 1500 python -m pytest tests/test_gpu_parity_scale.py tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -30 | tee gpurun_out/r02_run1_pytest_a.log
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -30 | tee gpurun_out/r02_run1_pytest_all.log
