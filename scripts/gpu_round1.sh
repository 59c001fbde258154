set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_4096.json 2> gpurun_out/bench_4096.err; tail -c 3000 gpurun_out/bench_4096.json; tail -5 gpurun_out/bench_4096.err
timeout 600 python bench.py --steps 100 --warmup 5 --envs-per-gpu 16384 --no-cpu-baseline > gpurun_out/bench_16384.json 2>> gpurun_out/bench_4096.err; tail -c 1500 gpurun_out/bench_16384.json
# launch list (cold-cache, serialised: compare SHARES)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 10 > gpurun_out/ncu_bench.log 2>&1
tail -25 gpurun_out/launches.csv
# full capture of the dominant kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 10 -c 2 -o gpurun_out/prof_env_step python bench.py --steps 15 --warmup 3 --no-cpu-baseline --e2e-steps 10 > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_act_kernel -s 10 -c 1 -o gpurun_out/prof_lstm_act python bench.py --steps 15 --warmup 3 --no-cpu-baseline --e2e-steps 10 > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out
