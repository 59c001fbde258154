#!/bin/bash
# compute-sanitizer over the hot-path kernels; logs -> gpurun_out/${TAG}_sanitizer_*.log (copied to profiles/)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=${1:-r02f}
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 python scripts/sanitize_target.py > gpurun_out/${TAG}_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok" gpurun_out/${TAG}_sanitizer_$tool.log | tail -12
done
IRRL_STEP_BLK=128 timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -c "
import sys; sys.path.insert(0, 'tests')
import numpy as np
from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg import trot_cfg
from gpu_lib import Cuda
c = Cuda(trot_cfg(num_envs=300)); c.reset()
for t in range(3): c.step(np.zeros((300, 12), np.float32))
print('barrier variant (128-thread CTAs) ok')" > gpurun_out/${TAG}_sanitizer_synccheck_blk128.log 2>&1; echo "synccheck blk128 rc=$?"; tail -3 gpurun_out/${TAG}_sanitizer_synccheck_blk128.log
