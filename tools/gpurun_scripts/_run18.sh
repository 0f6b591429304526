cd $GRAFT_REPO_ROOT
timeout 400 python tools/diag_determinism2.py 2048 100 7 2>&1 | grep DET2
