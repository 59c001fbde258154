#!/bin/bash
# device time of the learner kernels for build variants: bash scripts/r2c_variants.sh "<macros 1>" "<macros 2>" ...   ("" = default build)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for v in "$@"; do
  IRRL_EXP="$v" python -m high_speed_quadrupedal_locomotion_by_irrl_b200.build --force >/dev/null 2>&1 || { echo "build failed: $v"; continue; }
  echo "variant [$v]"
  python scripts/r2c_seq_ab.py 8192 750 2>&1 | grep -E "tensor-core  N=8192|autograd \[mma\]" | tail -4
  [ -n "$GEMM" ] && python scripts/r2c_gemm_ab.py 8192 750 2>&1 | grep -E "N=8192"
done
python -m high_speed_quadrupedal_locomotion_by_irrl_b200.build --force >/dev/null 2>&1
