#!/bin/bash
# A/B of build variants on the PPO update: usage  bash scripts/r2_ab_ppo.sh "<macros>" ...
cd $GRAFT_REPO_ROOT
for v in "$@"; do
  IRRL_EXP="$v" python -m high_speed_quadrupedal_locomotion_by_irrl_b200.build --force >/dev/null 2>&1 || { echo "build failed: $v"; continue; }
  python -m pytest tests/test_gpu_surface_and_rollout.py -q -m gpu -k "bptt or ppo_iteration" 2>&1 | tail -1
  python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline --e2e-steps 10 2>&1 | tail -1 | python -c "
import sys, json; d = json.loads(sys.stdin.readline())['ppo_iteration']; print('variant [$v]: iteration %.3f s  rollout %.3f  update %.3f' % (d['value'], d['rollout_s'], d['update_s']))"
done
