#!/bin/bash
# ncu --set full of the learner's tensor-core kernels at the learner's width (8192 envs), short T: sequence kernels and streaming products
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=${1:-r02c}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mma_kernel -s 6 -c 2 -o gpurun_out/${TAG}_lstm_seq_mma -f \
  python scripts/r2c_seq_ab.py 8192 96 > gpurun_out/${TAG}_ncu_seq.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rows_kernel -s 30 -c 17 -o gpurun_out/${TAG}_learner_gemm -f \
  python scripts/r2c_gemm_ab.py 8192 96 > gpurun_out/${TAG}_ncu_gemm.log 2>&1
for r in lstm_seq_mma learner_gemm; do
  ncu -i gpurun_out/${TAG}_$r.ncu-rep --page details > gpurun_out/${TAG}_${r}_details.txt 2>&1
  grep -E "^  [a-z_:A-Za-z<>0-9, ]+\(|Duration|Registers Per|Achieved Occupancy|Issued Ipc Active|No Eligible|Warp Cycles Per Issued|DRAM Throughput|Memory Throughput|highest-utilized" gpurun_out/${TAG}_${r}_details.txt | cut -c1-150
done
python scripts/ncu_source_summary.py gpurun_out/${TAG}_lstm_seq_mma.ncu-rep 30 0 > gpurun_out/${TAG}_lstm_seq_fwd_mma_source_summary.txt 2>&1
python scripts/ncu_source_summary.py gpurun_out/${TAG}_lstm_seq_mma.ncu-rep 30 1 > gpurun_out/${TAG}_lstm_seq_bwd_mma_source_summary.txt 2>&1
