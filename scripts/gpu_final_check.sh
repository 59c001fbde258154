#!/bin/bash
# end-of-round verification: full GPU suite, smoke, and a capture of the step kernel at the north-star per-GPU size (8192 robots)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 40 -c 1 -o gpurun_out/env_step_8192 -f python bench.py --envs-per-gpu 8192 --steps 50 --warmup 3 --no-cpu-baseline --e2e-steps 10 > gpurun_out/ncu_8k.log 2>&1
ncu -i gpurun_out/env_step_8192.ncu-rep --page details > gpurun_out/env_step_8192_details.txt 2>&1
tail -1 gpurun_out/ncu_8k.log | cut -c1-200
