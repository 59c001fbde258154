#!/bin/bash
# A/B of step-kernel build variants at 4096 and 16384 robots: usage  bash scripts/r2_ab.sh "<macros A>" "<macros B>" ...
cd $GRAFT_REPO_ROOT
for v in "$@"; do
  IRRL_EXP="$v" python -m high_speed_quadrupedal_locomotion_by_irrl_b200.build --force >/dev/null 2>&1 || { echo "build failed: $v"; continue; }
  for n in 4096 16384; do
    python bench.py --workload trot --envs-per-gpu $n --steps 300 --warmup 20 --no-extras --no-ppo --no-cpu-baseline --e2e-steps 10 2>&1 | tail -1 | python -c "
import sys, json; d = json.loads(sys.stdin.readline()); print('variant [$v] envs $n: step %.1f us  act %.1f us  value %.3e  clocks %s' % (1e3 * d['roofline']['kernel_ms'], 1e3 * d['roofline']['lstm_act']['kernel_ms'], d['value'], d['clocks']['sm_mhz']))"
  done
done
