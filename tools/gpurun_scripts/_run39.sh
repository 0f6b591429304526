cd $GRAFT_REPO_ROOT
timeout 100 python bench.py --config 1 --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg1_d.json 2> gpurun_out/r02_bench_cfg1_d.err
cut -c1-200 gpurun_out/r02_bench_cfg1_d.json
