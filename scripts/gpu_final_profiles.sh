cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
# 1. the bench line itself (default flags) + reference arm
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_default.json
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.json 2>> gpurun_out/bench_default.err
for n in 8192 16384 32768; do timeout 600 python bench.py --envs-per-gpu $n --no-cpu-baseline > gpurun_out/bench_$n.json 2>> gpurun_out/bench_default.err; done
timeout 600 python bench.py --workload relaxation --envs-per-gpu 16384 --no-cpu-baseline > gpurun_out/bench_relaxation_16384.json 2>> gpurun_out/bench_default.err
timeout 600 python bench.py --workload stairs --envs-per-gpu 32768 --no-cpu-baseline > gpurun_out/bench_stairs_32768.json 2>> gpurun_out/bench_default.err
# 2. launch list of the bench command (cold-cache, serialised: compare SHARES)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 80 --csv --log-file gpurun_out/r01_launches_final.csv python bench.py --steps 60 --warmup 5 --no-cpu-baseline --e2e-steps 10 > gpurun_out/ncu_l.log 2>&1
# 3. per-opcode counters (steady state)
M=smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__cycles_active.avg,sm__cycles_elapsed.max
timeout 900 ncu --metrics $M --clock-control none -k regex:"env_step_kernel|lstm_act" -s 300 -c 16 --csv --log-file gpurun_out/r01_counts_final.csv python bench.py --steps 160 --warmup 3 --no-cpu-baseline --e2e-steps 10 > gpurun_out/ncu_counts.log 2>&1
# 4. full captures
timeout 900 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 150 -c 1 -o gpurun_out/r01_env_step_final -f python bench.py --steps 160 --warmup 3 --no-cpu-baseline --e2e-steps 10 > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_act -s 150 -c 1 -o gpurun_out/r01_lstm_act_final -f python bench.py --steps 160 --warmup 3 --no-cpu-baseline --e2e-steps 10 > gpurun_out/ncu_full2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lstm_seq_fwd_kernel|lstm_seq_bwd_kernel" -c 2 -o gpurun_out/r01_lstm_seq_final -f python scripts/ppo_timing.py --envs 2048 --steps 750 --epochs 1 --iters 1 > gpurun_out/ncu_full3.log 2>&1
for k in env_step lstm_act lstm_seq; do ncu -i gpurun_out/r01_${k}_final.ncu-rep --page details > gpurun_out/r01_${k}_final_details.txt 2>&1; done
ls -la gpurun_out | tail -20
