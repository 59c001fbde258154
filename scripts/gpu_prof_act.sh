cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-4096}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_act_kernel -s 150 -c 1 -o gpurun_out/prof_lstm_act_$N -f python bench.py --steps 160 --warmup 3 --no-cpu-baseline --e2e-steps 10 --envs-per-gpu $N > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
