cd $GRAFT_REPO_ROOT
make -C oracle >/dev/null 2>&1
timeout 600 python -m pytest tests/test_gpu_env_parity.py -x -q -m gpu 2>&1 | tail -2
run() { for n in 4096 8192 16384 32768; do timeout 300 python bench.py --steps 150 --warmup 10 --envs-per-gpu $n --no-cpu-baseline --e2e-steps 20 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('  N=%d value=%.3e ms/step=%.4f step_kernel_ms=%.4f act_kernel_ms=%.4f frac=%.3f' % (d['config']['envs_per_gpu'], d['value'], d['ms_per_step'], r['kernel_ms'], r['lstm_act']['kernel_ms'], r['frac'] or 0))
"; done; }
echo "== default (minblocks 1)"; run
for mb in 6 8; do echo "== minblocks $mb"; IRRL_STEP_MINBLOCKS=$mb python -m high_speed_quadrupedal_locomotion_by_irrl_b200.build --force >/dev/null 2>&1; run; done
