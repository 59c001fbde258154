#!/bin/bash
# ncu --set full of the tensor-core sequence kernels (fwd + bwd) at the learner's width, short T; then A/B of build variants by device time.
# usage: bash scripts/r2c_seq_prof.sh TAG "<macros of variant 1>" "<macros of variant 2>" ...
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=${1:-r02c}; shift
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mma_kernel -s 6 -c 2 -o gpurun_out/${TAG}_lstm_seq_mma -f \
  python scripts/r2c_seq_ab.py 8192 96 > gpurun_out/${TAG}_ncu_seq.log 2>&1
ncu -i gpurun_out/${TAG}_lstm_seq_mma.ncu-rep --page details > gpurun_out/${TAG}_lstm_seq_mma_details.txt 2>&1
python scripts/ncu_source_summary.py gpurun_out/${TAG}_lstm_seq_mma.ncu-rep > gpurun_out/${TAG}_lstm_seq_mma_source_summary.txt 2>&1
grep -E "lstm_seq|Duration|Registers Per|Achieved Occupancy|Issued Ipc|No Eligible|Warp Cycles Per Issued|Local|Tensor|DRAM Throughput" gpurun_out/${TAG}_lstm_seq_mma_details.txt | head -40
for v in "$@"; do
  IRRL_EXP="$v" python -m high_speed_quadrupedal_locomotion_by_irrl_b200.build --force >/dev/null 2>&1 || { echo "build failed: $v"; continue; }
  echo "variant [$v]"; python scripts/r2c_seq_ab.py 8192 750 2>&1 | grep -E "N=8192|autograd \[mma\]" | tail -4
done
