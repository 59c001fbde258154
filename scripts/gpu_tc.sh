#!/bin/bash
# bring-up of the tcgen05 act kernel on the GPU box; each stage in its own process with its own timeout
mkdir -p gpurun_out
for s in probe act time; do
  timeout 180 python tests/tools/tc_bringup.py $s > gpurun_out/tc_$s.log 2>&1; echo "stage $s rc=$?" >> gpurun_out/tc_$s.log
  cat gpurun_out/tc_$s.log
done
if [ "$1" = "ncu" ]; then
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:lstm_act_tc -s 20 -c 1 -o gpurun_out/tc_act -f python tests/tools/tc_bringup.py time > gpurun_out/tc_ncu.log 2>&1
  ncu -i gpurun_out/tc_act.ncu-rep --page details > gpurun_out/tc_act_details.txt 2>&1
  tail -3 gpurun_out/tc_ncu.log
fi
