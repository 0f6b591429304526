cd $GRAFT_REPO_ROOT
( timeout 10 python examples/dmrg.py sweep_loop_version 2>&1 | tail -3 ) 
( timeout 10 python examples/timed_dmrg.py --N 20 2>&1 | tail -7 )
( timeout 10 python examples/quench_evolution.py --N 6 --total-time 0.2 2>&1 | tail -4 )
( timeout 10 python examples/tdvp.py test_tdvp --total-time 0.1 2>&1 | tail -5 )
( timeout 10 python examples/fitting.py 2>&1 | tail -6 )
( timeout 10 python examples/dmrg.py dmrg 2>&1 | tail -4 )
( timeout 10 python examples/tdvp.py tdvp --N 8 --total-time 0.2 2>&1 | tail -6 )
