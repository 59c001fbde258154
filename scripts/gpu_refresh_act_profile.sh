#!/bin/bash
# refresh the act-kernel evidence after the last kernel change (counts, launch list, full capture)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 80 --csv --log-file gpurun_out/r01_launches_final.csv python bench.py --steps 60 --warmup 5 --no-cpu-baseline --e2e-steps 10 > gpurun_out/ncu_l.log 2>&1
M=smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__cycles_active.avg,sm__cycles_elapsed.max
timeout 900 ncu --metrics $M --clock-control none -k regex:"env_step_kernel|lstm_act" -s 300 -c 16 --csv --log-file gpurun_out/r01_counts_final.csv python bench.py --steps 160 --warmup 3 --no-cpu-baseline --e2e-steps 10 > gpurun_out/ncu_counts.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_act -s 150 -c 1 -o gpurun_out/r01_lstm_act_final -f python bench.py --steps 160 --warmup 3 --no-cpu-baseline --e2e-steps 10 > gpurun_out/ncu_full2.log 2>&1
ncu -i gpurun_out/r01_lstm_act_final.ncu-rep --page details > gpurun_out/r01_lstm_act_final_details.txt 2>&1
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 300 gpurun_out/bench_default.json
