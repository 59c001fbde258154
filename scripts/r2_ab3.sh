#!/bin/bash
# A/B of build-time environment knobs (build.py reads IRRL_*): usage  bash scripts/r2_ab3.sh "IRRL_STEP_MINWARPS=12" ...
cd $GRAFT_REPO_ROOT
for v in "$@"; do
  env $v python -m high_speed_quadrupedal_locomotion_by_irrl_b200.build --force -v 2>&1 | grep -A3 "env_step_kernelILi64" | grep -E "stack|Used"
  for n in 4096 16384 32768; do
    python bench.py --workload trot --envs-per-gpu $n --steps 300 --warmup 20 --no-extras --no-ppo --no-cpu-baseline --e2e-steps 10 2>&1 | tail -1 | python -c "
import sys, json; d = json.loads(sys.stdin.readline()); print('build [$v] envs $n: step %.1f us  act %.1f us  value %.3e' % (1e3 * d['roofline']['kernel_ms'], 1e3 * d['roofline']['lstm_act']['kernel_ms'], d['value']))"
  done
done
