set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_hotpath.py tests/test_gpu_multi.py -x -q -m gpu -k "two_gpus or slab or multi" 2>&1 | tail -15 | tee gpurun_out/r02_run26_pytest.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --chi 2048 --nsites 40 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2_chi2048_b.json 2> gpurun_out/r02_bench_n2_chi2048_b.err
grep -E "sharded-vs-single|Error|error" gpurun_out/r02_bench_n2_chi2048_b.err | head -5
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n2_chi2048_b.json').read().strip().splitlines()[-1])
for k in ['value','ms_per_step','e2e','sharded_vs_replicated_max_rel_err','sharded_parity_ok','region_step_s','region_phase_ms','hbm_pool_used_gib','env_hbm_gib_per_gpu','env_hbm_gib_if_replicated']: print(k, d.get(k))
PY
