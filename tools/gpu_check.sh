#!/bin/bash
# Runs the GPU test-suite and collects logs under gpurun_out/.  Usage: tools/gpu_check.sh [pytest args]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout=600 "$@" 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
