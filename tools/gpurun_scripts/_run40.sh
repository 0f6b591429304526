cd $GRAFT_REPO_ROOT
timeout 45 python bench.py --chi 1024 --nsites 28 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_chi1024_sweepcheck.json 2> gpurun_out/r02_bench_chi1024_sweepcheck.err
tail -c 300 gpurun_out/r02_bench_chi1024_sweepcheck.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_chi1024_sweepcheck.json').read().strip().splitlines()[-1])
for k in ['value','ms_per_step','region_step_s','full_sweep_s','full_sweep_energy','full_sweep_maxlinkdim','full_sweep_launches']: print(k, d.get(k))
PY
