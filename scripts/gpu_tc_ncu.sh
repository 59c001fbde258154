#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on --warp-sampling-interval 0 -k regex:lstm_act_tc -s 20 -c 1 -o gpurun_out/tc_act -f python tests/tools/tc_bringup.py time > gpurun_out/tc_ncu.log 2>&1
ncu -i gpurun_out/tc_act.ncu-rep --page details > gpurun_out/tc_act_details.txt 2>&1
tail -3 gpurun_out/tc_ncu.log
