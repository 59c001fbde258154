#!/bin/bash
# A/B of build-time and run-time knobs of the step kernel.  usage: bash scripts/r2_ab4.sh "<build env>|<run env>|<sizes>" ...
# e.g. "IRRL_EXP=EXP_X|IRRL_STEP_BLK=256|4096 16384"
cd $GRAFT_REPO_ROOT
last_build="__none__"
for v in "$@"; do
  IFS='|' read -r benv renv sizes <<< "$v"
  [ -z "$sizes" ] && sizes="4096 16384 32768"
  if [ "$benv" != "$last_build" ]; then
    env $benv python -m high_speed_quadrupedal_locomotion_by_irrl_b200.build --force -v 2>&1 | grep -A2 "env_step_kernelILi128ELb1ELb0" | grep -E "stack|Used" | tr '\n' ' '; echo
    last_build="$benv"
  fi
  for n in $sizes; do
    env $renv python bench.py --workload trot --envs-per-gpu $n --steps 300 --warmup 20 --no-extras --no-ppo --no-cpu-baseline --e2e-steps 10 2>&1 | tail -1 | python -c "
import sys, json; d = json.loads(sys.stdin.readline()); print('[$benv|$renv] envs $n: step %.1f us  act %.1f us  value %.3e' % (1e3 * d['roofline']['kernel_ms'], 1e3 * d['roofline']['lstm_act']['kernel_ms'], d['value']))"
  done
done
