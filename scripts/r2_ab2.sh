#!/bin/bash
# A/B of run-time knobs of the step kernel: usage  bash scripts/r2_ab2.sh "ENV=VAL ..." ...
cd $GRAFT_REPO_ROOT
for v in "$@"; do
  for n in 4096 16384; do
    env $v python bench.py --workload trot --envs-per-gpu $n --steps 300 --warmup 20 --no-extras --no-ppo --no-cpu-baseline --e2e-steps 10 2>&1 | tail -1 | python -c "
import sys, json; d = json.loads(sys.stdin.readline()); print('knobs [$v] envs $n: step %.1f us  act %.1f us  value %.3e' % (1e3 * d['roofline']['kernel_ms'], 1e3 * d['roofline']['lstm_act']['kernel_ms'], d['value']))"
  done
done
