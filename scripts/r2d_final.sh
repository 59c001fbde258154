#!/bin/bash
# final round-2 verification of the learner work: GPU suite, smoke, sanitizers, ncu of the tcgen05 learner kernels, default bench line
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash scripts/r2_sanitize.sh r02d 2>&1 | grep -E "rc=|SUMMARY" | head -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_kernel -s 18 -c 20 -o gpurun_out/r02d_learner_tc -f \
  python scripts/r2c_gemm_ab.py 8192 96 > gpurun_out/r02d_ncu_tc.log 2>&1
ncu -i gpurun_out/r02d_learner_tc.ncu-rep --page details > gpurun_out/r02d_learner_tc_details.txt 2>&1
grep -E "^  [a-z_:A-Za-z<>0-9, ]+\(|Duration|Registers Per|Issued Ipc Active|DRAM Throughput|highest-utilized" gpurun_out/r02d_learner_tc_details.txt | cut -c1-140 | head -40
python bench.py 2>&1 | tail -1 > gpurun_out/r02d_bench_default.json; cut -c1-300 gpurun_out/r02d_bench_default.json
