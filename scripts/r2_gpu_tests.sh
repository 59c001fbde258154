#!/bin/bash
# round 2: GPU parity suite (+ optional bench line).  usage: gpurun -- bash scripts/r2_gpu_tests.sh [pytest-args...]
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
t0=$(date +%s)
timeout 1500 python -m pytest tests -q -m gpu -x "$@" 2>&1 | tail -25 | tee gpurun_out/r2_pytest_gpu.log
echo "pytest seconds: $(( $(date +%s) - t0 ))"
