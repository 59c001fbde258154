#!/bin/bash
# 30 PPO iterations from random weights at 2048 envs through the tensor-core learner: the mean episode reward must rise (functional check of the whole loop)
cd $GRAFT_REPO_ROOT
python - <<'PY'
import yaml, sys
sys.path.insert(0,'.')
from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg import trot_cfg
yaml.safe_dump({"seed":1,"record_video":False,"environment":trot_cfg(num_envs=2048, StochasticDynamics=True, ObsNoise=2.0)}, open("/tmp/trot.yaml","w"), sort_keys=False)
PY
timeout 900 python scripts/run_bp_v5.py --cfg /tmp/trot.yaml --max_iter 46080000 --l 1e-3 --save false 2>&1 | grep -E "nupdates|Error|error" | awk '{print $2, $3, $4, $5, $6, $7, $8, $9, $10, $11, $12, $13, $14, $15, $16, $17, $18, $19, $20, $21, $22, $23, $24, $25, $26, $27}' | cut -c1-230 | awk 'NR<=2 || NR%5==0' | tail -10
