#!/usr/bin/env python
"""Training driver with the reference's CLI for the hot path's caller (IRRL/script/run_bp_v5.py:29-108, 196-259):
`--train` (default) / `--load <pkl|npz>` (relaxation: re-run with other reward coefficients in the YAML, readme.md:71-78)
`--cfg`, `--l`, `--max_iter`, `--save`.  The interactive `--test` tele-op branch is out of scope.

    python scripts/run_bp_v5.py --cfg flex_gym/env/env/BlackPanther_V55/urdf/default_cfg.yaml --max_iter 3000000 --num_envs 4096
    torchrun --nproc-per-node 8 scripts/run_bp_v5.py --num_envs 65536 ...      # envs sharded, gradients all-reduced
"""
import argparse
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def parse_args(argv):
    from flex_gym.env.env.BlackPanther_V55 import __BLACKPANTHER_V55_RESOURCE_DIRECTORY__ as __RSCDIR__
    p = argparse.ArgumentParser(description="Train control policies.")
    p.add_argument("--train", dest="train", action="store_true", default=True)
    p.add_argument("--cfg", type=str, default=os.path.abspath(__RSCDIR__ + "/default_cfg.yaml"), help="configuration file")
    p.add_argument("--max_iter", dest="max_iter", type=int, default=200000000, help="total env steps")
    p.add_argument("--save", dest="save_flag", type=lambda s: s.lower() not in ("0", "false"), default=True)
    p.add_argument("--l", dest="learn_rate", type=float, default=1e-3)
    p.add_argument("--load", dest="pre_trained_model", type=str, default=None, help="pre-trained model (.pkl of the reference or .npz)")
    p.add_argument("--num_envs", type=int, default=None, help="override environment.num_envs (total over all ranks)")
    p.add_argument("--log_dir", type=str, default=os.path.join(ROOT, "data", "black_panther_v5"))
    return p.parse_args(argv)


def main(argv=None):
    args = parse_args(argv if argv is not None else sys.argv[1:])
    import torch
    from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg import load_yaml_file
    from high_speed_quadrupedal_locomotion_by_irrl_b200.sharding import env_from_torchrun, make_sharded_env
    from flex_gym.env.RaisimGymVecEnv import RaisimGymVecEnv as Environment
    from flex_gym.env.env.BlackPanther_V55 import __BLACKPANTHER_V55_RESOURCE_DIRECTORY__ as __RSCDIR__
    from flex_gym.algo.ppo2 import PPO2
    from flex_gym.helper.raisim_gym_helper import ConfigurationSaver
    rank, local, world = env_from_torchrun()
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = load_yaml_file(args.cfg)
    envcfg = dict(cfg["environment"]); envcfg["render"] = False
    total = args.num_envs or int(envcfg["num_envs"])
    env = Environment(make_sharded_env(envcfg, total, __RSCDIR__))
    saver = ConfigurationSaver(log_dir=args.log_dir, save_items=[args.cfg]) if (args.save_flag and rank == 0) else None
    kw = dict(gamma=0.99, n_steps=math.floor(envcfg["max_time"] / envcfg["control_dt"]), ent_coef=0.0, learning_rate=args.learn_rate, vf_coef=0.5,
              max_grad_norm=0.5, lam=0.998, nminibatches=1, noptepochs=10, cliprange=0.2, verbose=1)                     # run_bp_v5.py:227-242
    model = PPO2.load(args.pre_trained_model, env, **kw) if args.pre_trained_model else PPO2(env, **kw)                  # run_bp_v5.py:244-249
    model.learn(total_timesteps=args.max_iter, log_dir=(saver.data_dir + "/model") if saver else None)
    if saver and rank == 0:
        model.save(saver.data_dir + "/final")
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
