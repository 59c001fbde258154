#!/bin/bash
# step-kernel CTA size / phase-barrier variants (rebuilt on the box), step kernel time at several env counts
cd $GRAFT_REPO_ROOT
run() { for n in 4096 8192 32768; do timeout 300 python bench.py --steps 100 --warmup 10 --envs-per-gpu $n --no-cpu-baseline --e2e-steps 10 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('  N=%d value=%.3e ms/step=%.4f step_kernel_ms=%.4f act_kernel_ms=%.4f frac=%.3f' % (d['config']['envs_per_gpu'], d['value'], d['ms_per_step'], r['kernel_ms'], r['lstm_act']['kernel_ms'], r['frac'] or 0))
"; done; }
for v in "64 0" "64 1" "128 1" "256 1" "256 0" "128 0"; do set -- $v; echo "== block $1 sync $2"; IRRL_STEP_BLOCK=$1 IRRL_STEP_SYNC=$2 python -m high_speed_quadrupedal_locomotion_by_irrl_b200.build --force >/dev/null 2>&1; run; done
