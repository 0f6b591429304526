cd $GRAFT_REPO_ROOT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
grep -E "sharded-vs-single|Error|error|Traceback" gpurun_out/r02_bench_n2.err | head -5
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n2.json').read().strip().splitlines()[-1])
for k in ['value','ms_per_step','e2e','sharded_vs_replicated_max_rel_err','sharded_parity_ok','region_step_s','region_phase_ms','hbm_pool_used_gib','env_hbm_gib_per_gpu','env_hbm_gib_if_replicated','sweep_s_extrapolated']: print(k, d.get(k))
PY
