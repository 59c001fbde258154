cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err
grep '^{' gpurun_out/bench_${n}gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('n_gpus', d['n_gpus'], 'value %.4e'%d['value'], 'e2e %.4e'%d['e2e']['value'], 'ms/step', d['ms_per_step'], d['config']['total_envs'])"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 200 --warmup 10 --no-cpu-baseline --envs-per-gpu 8192 > gpurun_out/bench_8gpu_65536.json 2> gpurun_out/bench_8gpu_65536.err
grep '^{' gpurun_out/bench_8gpu_65536.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('n_gpus', d['n_gpus'], 'value %.4e'%d['value'], 'e2e %.4e'%d['e2e']['value'], 'ms/step', d['ms_per_step'], d['config']['total_envs'])"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 scripts/ppo_timing.py --envs 8192 --steps 750 --epochs 10 --iters 2 2>&1 | tail -1
