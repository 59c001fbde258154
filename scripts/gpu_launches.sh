cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-4096}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"env_step_kernel|lstm_act_kernel" -s 300 -c 8 --csv --log-file gpurun_out/launches_$N.csv python bench.py --steps 160 --warmup 3 --no-cpu-baseline --e2e-steps 10 --envs-per-gpu $N > gpurun_out/ncu_l.log 2>&1
awk -F'","' '{print $5, $(NF)}' gpurun_out/launches_$N.csv | tail -8
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"env_step_kernel|lstm_act_kernel" -s 300 -c 4 --csv --log-file gpurun_out/launches_nf_$N.csv python bench.py --steps 160 --warmup 3 --no-cpu-baseline --e2e-steps 10 --envs-per-gpu $N --no-l2-flush > gpurun_out/ncu_l.log 2>&1
awk -F'","' '{print $5, $(NF)}' gpurun_out/launches_nf_$N.csv | tail -4
