cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -c 1200 gpurun_out/bench_2gpu.json; tail -3 gpurun_out/bench_2gpu.err
timeout 600 python bench.py --impl reference --steps 100 --warmup 5 | tail -c 900
timeout 900 python scripts/ppo_timing.py --envs 1024 --steps 750 --epochs 2 --iters 2 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/ppo_timing.py --envs 1024 --steps 750 --epochs 2 --iters 2 2>&1 | tail -3
