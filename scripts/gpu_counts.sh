cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-4096}
M=smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__cycles_active.avg,sm__cycles_elapsed.max
timeout 900 ncu --metrics $M --clock-control none -k regex:"env_step_kernel|lstm_act_kernel" -s 300 -c 16 --csv --log-file gpurun_out/counts_$N.csv python bench.py --steps 160 --warmup 3 --no-cpu-baseline --e2e-steps 10 --envs-per-gpu $N > gpurun_out/ncu_counts.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 150 -c 1 -o gpurun_out/prof_env_step_$N -f python bench.py --steps 160 --warmup 3 --no-cpu-baseline --e2e-steps 10 --envs-per-gpu $N > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/launches_$N.csv python bench.py --steps 40 --warmup 3 --no-cpu-baseline --e2e-steps 10 --envs-per-gpu $N > gpurun_out/ncu_bench.log 2>&1
ls gpurun_out | head -30
