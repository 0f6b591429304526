cd $GRAFT_REPO_ROOT
python tools/diag_determinism.py 1024 100 2>&1 | grep DET
python tools/diag_determinism.py 1024 100 qr_block_min=0 2>&1 | grep DET
timeout 900 python -m pytest tests/test_gpu_shard_emulation.py -x -q -m gpu 2>&1 | tail -8
