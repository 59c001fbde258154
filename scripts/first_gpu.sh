set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
cd $GRAFT_REPO_ROOT
make -C oracle 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_env_parity.py -x -q -m gpu 2>&1 | tail -40
