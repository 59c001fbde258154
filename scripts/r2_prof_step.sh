#!/bin/bash
# ncu --set full capture of one env_step_kernel launch in the contact-rich steady state (launch 150 of the bench), N envs (default 16384)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=${1:-16384}; TAG=${2:-r02}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 150 -c 1 -o gpurun_out/${TAG}_env_step_$N -f \
  python bench.py --steps 160 --warmup 3 --no-cpu-baseline --no-extras --no-ppo --e2e-steps 10 --envs-per-gpu $N > gpurun_out/${TAG}_ncu_step_$N.log 2>&1
ncu -i gpurun_out/${TAG}_env_step_$N.ncu-rep --page details > gpurun_out/${TAG}_env_step_${N}_details.txt 2>&1
python scripts/ncu_source_summary.py gpurun_out/${TAG}_env_step_$N.ncu-rep > gpurun_out/${TAG}_env_step_${N}_source_summary.txt 2>&1
grep -E "Duration|Registers Per|Achieved Occupancy|Issued Ipc|No Eligible|Warp Cycles Per Issued|Local" gpurun_out/${TAG}_env_step_${N}_details.txt | head -12
sed -n 1,12p gpurun_out/${TAG}_env_step_${N}_source_summary.txt; grep "stall reasons" gpurun_out/${TAG}_env_step_${N}_source_summary.txt
