cd $GRAFT_REPO_ROOT
python tools/diag_emul.py 2048 100 2>&1 | grep EMUL
python tools/diag_emul.py 2048 40 2>&1 | grep EMUL
timeout 900 python -m pytest tests/test_gpu_shard_emulation.py -x -q -m gpu 2>&1 | tail -15
