set -x
cd $GRAFT_REPO_ROOT
B="python bench.py --chi 4096 --nsites 26 --steps 3 --warmup 3 --no-full-sweep --no-cpu-baseline"
NSB_PROFILE_TIMED=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_matvec_chi4096.csv $B --no-region-step > gpurun_out/ncu_a.log 2>&1
NSB_PROFILE_TIMED=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tma -c 3 -f -o gpurun_out/r02_prof_gemm $B --no-region-step > gpurun_out/ncu_b.log 2>&1
NSB_PROFILE_REGION=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_ncu_launches_region_chi4096.csv $B --region-steps 3 > gpurun_out/ncu_c.log 2>&1
NSB_PROFILE_TIMED=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_cfg3.csv python bench.py --config 3 --steps 3 --warmup 3 --no-cpu-baseline --region-steps 1 > gpurun_out/ncu_d.log 2>&1
NSB_PROFILE_TIMED=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_grouped -c 3 -f -o gpurun_out/r02_prof_grouped python bench.py --config 3 --steps 3 --warmup 3 --no-cpu-baseline --region-steps 1 > gpurun_out/ncu_e.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv
tail -3 gpurun_out/ncu_a.log gpurun_out/ncu_b.log gpurun_out/ncu_c.log gpurun_out/ncu_d.log gpurun_out/ncu_e.log
