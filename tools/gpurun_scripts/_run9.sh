set -x
cd $GRAFT_REPO_ROOT
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$T --master-port 29620 tools/diag_nccl_fp64.py 2>&1 | grep NCCL_FP64
NCCL_NVLS_ENABLE=0 $T --master-port 29621 tools/diag_nccl_fp64.py 2>&1 | grep NCCL_FP64
NCCL_ALGO=Ring $T --master-port 29622 tools/diag_nccl_fp64.py 2>&1 | grep NCCL_FP64
NCCL_NVLS_ENABLE=0 timeout 900 $T --master-port 29623 bench.py --gpus 8 --steps 10 --warmup 3 --chi 2048 --nsites 40 --region-steps 2 > gpurun_out/r02_bench_n8_chi2048_nonvls.json 2> gpurun_out/r02_bench_n8_chi2048_nonvls.err
grep -n "AssertionError" gpurun_out/r02_bench_n8_chi2048_nonvls.err | head -2
cut -c1-400 gpurun_out/r02_bench_n8_chi2048_nonvls.json
timeout 900 $T --master-port 29624 bench.py --gpus 8 --steps 10 --warmup 3 --chi 2048 --nsites 40 --region-steps 2 > gpurun_out/r02_bench_n8_chi2048.json 2> gpurun_out/r02_bench_n8_chi2048.err
grep -n "AssertionError" gpurun_out/r02_bench_n8_chi2048.err | head -2
cut -c1-400 gpurun_out/r02_bench_n8_chi2048.json
