cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for n in 4096 8192 16384; do
timeout 300 python bench.py --steps 200 --warmup 10 --envs-per-gpu $n --no-cpu-baseline --e2e-steps 30 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('N=%d value=%.3e e2e=%.3e ms/step=%.4f step_kernel_ms=%.4f act_kernel_ms=%.4f frac=%s clocks=%s rew=%.3f eps=%d' % (d['config']['envs_per_gpu'], d['value'], d['e2e']['value'], d['ms_per_step'], r['kernel_ms'], r['lstm_act']['kernel_ms'], r['frac'], d['clocks']['sm_mhz'], d['sanity']['mean_reward'], d['sanity']['episodes_finished']), d['sanity']['gs_sweeps_last_substep']['mean'])
    else: print(l.rstrip()[:300])
"
done
