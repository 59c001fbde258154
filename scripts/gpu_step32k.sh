#!/bin/bash
# env_step_kernel at 32768 envs: full ncu capture (stall reasons at high occupancy)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 40 -c 1 -o gpurun_out/env_step_32768 -f python bench.py --envs-per-gpu 32768 --steps 50 --warmup 3 --no-cpu-baseline --e2e-steps 10 > gpurun_out/ncu_32k.log 2>&1
ncu -i gpurun_out/env_step_32768.ncu-rep --page details > gpurun_out/env_step_32768_details.txt 2>&1
tail -2 gpurun_out/ncu_32k.log
