#!/bin/bash
# round 2 final evidence: launch list of the bench command, --set full captures of the step kernel (4096 / 16384 robots) and of the
# tcgen05 act kernel (16384), summaries written next to them.  usage: gpurun -- bash scripts/r2_final_profiles.sh [TAG]
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=${1:-r02f}
BENCH="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extras --no-ppo --e2e-steps 10"
# 1. launch list (gpu__time_duration.sum) of the bench command: the kernels' SHARE of a step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 120 --csv --log-file gpurun_out/${TAG}_launches_bench16384.csv $BENCH > gpurun_out/${TAG}_ncu_launches.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/${TAG}_launches_bench16384.csv")) if len(r) > 10 and r[0].isdigit()]
t = collections.defaultdict(list)
for r in rows: t[r[4].split("(")[0][:60]].append(float(r[-1]))
tot = sum(sum(v) for v in t.values())
for k, v in sorted(t.items(), key=lambda kv: -sum(kv[1])): print("%-62s n=%3d  mean %8.1f us  share %5.1f %%" % (k, len(v), sum(v) / len(v) / (1e3 if max(v) > 5e3 else 1), 100 * sum(v) / tot))
PY
# 2. step kernel, 4096 and 16384 robots
for N in 4096 16384; do bash scripts/r2_prof_step.sh $N $TAG; python scripts/ncu_line_summary.py gpurun_out/${TAG}_env_step_$N.ncu-rep 40 > gpurun_out/${TAG}_env_step_${N}_line_summary.txt 2>&1; done
# 3. act kernel at 16384
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_act_tc_kernel -s 30 -c 1 -o gpurun_out/${TAG}_lstm_act_tc_16384 -f $BENCH > gpurun_out/${TAG}_ncu_act.log 2>&1
ncu -i gpurun_out/${TAG}_lstm_act_tc_16384.ncu-rep --page details > gpurun_out/${TAG}_lstm_act_tc_16384_details.txt 2>&1
python scripts/ncu_source_summary.py gpurun_out/${TAG}_lstm_act_tc_16384.ncu-rep > gpurun_out/${TAG}_lstm_act_tc_16384_source_summary.txt 2>&1
grep -E "Duration|Registers Per|Achieved Occupancy" gpurun_out/${TAG}_lstm_act_tc_16384_details.txt | head -4
ls -la gpurun_out/${TAG}_*
