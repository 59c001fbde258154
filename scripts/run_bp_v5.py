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
    # headless subset of the reference's test mode (run_bp_v5.py:261-470: gamepad, OGRE and matplotlib parts are out of scope)
    p.add_argument("--test", dest="train", action="store_false", help="run a trained model in the Manual test configuration")
    p.add_argument("--model", dest="trained_model", type=str, default=None, help="trained model for --test (.pkl of the reference or .npz)")
    p.add_argument("--fix_cmd", dest="flag_fix_cmd", type=float, default=None, help="forward-speed command [m/s] instead of the gamepad (run_bp_v5.py:72)")
    p.add_argument("--eval", dest="flag_eval", action="store_true", default=False, help="print speed / height / stride statistics (run_bp_v5.py:58)")
    p.add_argument("--o", dest="flag_output", action="store_true", default=False, help="export the pi tower as CSV like CustomerLstmNN.save_model (run_bp_v5.py:78)")
    p.add_argument("--save_data", type=str, default="NoSave", help="npz file for OriginState / joint effort / generalised force / reward records (run_bp_v5.py:96)")
    p.add_argument("--seconds", type=float, default=8.0, help="length of the test run")
    p.add_argument("--cfg", type=str, default=os.path.abspath(__RSCDIR__ + "/default_cfg.yaml"), help="configuration file")
    p.add_argument("--max_iter", dest="max_iter", type=int, default=200000000, help="total env steps")
    p.add_argument("--save", dest="save_flag", type=lambda s: s.lower() not in ("0", "false"), default=True)
    p.add_argument("--l", dest="learn_rate", type=float, default=1e-3)
    p.add_argument("--load", dest="pre_trained_model", type=str, default=None, help="pre-trained model (.pkl of the reference or .npz)")
    p.add_argument("--num_envs", type=int, default=None, help="override environment.num_envs (total over all ranks)")
    p.add_argument("--log_dir", type=str, default=os.path.join(ROOT, "data", "black_panther_v5"))
    return p.parse_args(argv)


def test_mode(args):
    """run_bp_v5.py:261-470 without the gamepad: the command is --fix_cmd (low-pass filtered like GaitGenerator's command filter),
    written into the network input (Manual: True keeps the env's own command at zero, ENV:1013-1019); friction 0.8 (run_bp_v5.py:317)."""
    import json
    import numpy as np
    from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg import test_cfg, load_yaml_file, dump_yaml
    from high_speed_quadrupedal_locomotion_by_irrl_b200._flexible_robot import FlexibleGymEnv
    from high_speed_quadrupedal_locomotion_by_irrl_b200.vec_env import RaisimGymVecEnv
    from high_speed_quadrupedal_locomotion_by_irrl_b200.policy import FusedLstmPolicy, PARAM_NAMES, load_reference_params
    if args.trained_model is None:
        print("model path can't be ignored during test mode"); print("run -h to help"); sys.exit(1)          # run_bp_v5.py:267-270
    if args.trained_model.endswith(".npz"):
        z = np.load(args.trained_model); W = [z[k] for k in PARAM_NAMES]
    else:
        W = load_reference_params(args.trained_model)
    if args.cfg and os.path.exists(args.cfg) and "default_cfg" not in os.path.basename(args.cfg):
        envcfg = dict(load_yaml_file(args.cfg)["environment"]); envcfg["render"] = False
    else:
        envcfg = test_cfg(num_envs=1, render=False)
    n = int(envcfg["num_envs"]); dt = float(envcfg["control_dt"])
    env = RaisimGymVecEnv(FlexibleGymEnv("", dump_yaml(envcfg)))
    env.SetContactCoefficient(np.tile(np.array([[0.8, 0.2, 0.01]], np.float32), (n, 1)))
    pol = FusedLstmPolicy(W, n_env=n)
    if args.flag_output:
        from high_speed_quadrupedal_locomotion_by_irrl_b200.ppo2 import LstmActorCritic
        out_dir = os.path.join(ROOT, "data", "model_csv", os.path.splitext(os.path.basename(args.trained_model))[0])
        os.makedirs(out_dir, exist_ok=True)
        names = {"lstm_pi0_wx": "lstm1_wx", "lstm_pi0_wh": "lstm1_wh", "lstm_pi0_b": "lstm1_b", "lstm_pi1_wx": "lstm2_wx", "lstm_pi1_wh": "lstm2_wh",
                 "lstm_pi1_b": "lstm2_b", "pi_w": "pi_w", "pi_b": "pi_b"}
        for k, w in zip(PARAM_NAMES, W):
            if k in names:
                np.savetxt(os.path.join(out_dir, names[k] + ".csv"), np.asarray(w).reshape(w.shape[0], -1) if np.ndim(w) > 1 else np.asarray(w)[None], delimiter=",")
        print("pi tower exported to", out_dir)
    obs = env.reset(); state = np.zeros((n, 384), np.float32); done = np.zeros(n, bool)
    T = int(args.seconds / dt); cmd = 0.0; target = 0.0 if args.flag_fix_cmd is None else float(args.flag_fix_cmd)
    vx_half = float(envcfg["Vx"]) / 2.0
    rec = dict(origin=np.zeros((T, n, 41), np.float32), effort=np.zeros((T, n, 12), np.float32), gforce=np.zeros((T, n, 18), np.float32),
               reward=np.zeros((T, n), np.float32), extra=np.zeros((T, n, 6), np.float32), cmd=np.zeros(T, np.float32))
    falls = 0
    for t in range(T):
        cmd = 0.999 * cmd + 0.001 * target
        obs[:, 0] = (cmd - vx_half) / 1.0; obs[:, 1] = 0.0; obs[:, 2] = 0.0                      # obMean / obStd of the command slots (ENV:375-390)
        act, val, state, nlp = pol.step(obs, state, done, deterministic=True)
        obs, rew, done, info = env.step(np.clip(act, -1, 1))
        falls += int(done.sum())
        rec["origin"][t] = env.OriginState(); rec["effort"][t] = env.GetJointEffort(); rec["gforce"][t] = env.GetGeneralizedForce()
        rec["reward"][t] = rew; rec["extra"][t] = info._extra; rec["cmd"][t] = cmd
    if args.save_data != "NoSave":
        np.savez_compressed(args.save_data, **rec, control_dt=dt)
        print("saved", args.save_data)
    if args.flag_eval:
        half = rec["origin"][T // 2:]
        sig = half[:, 0, 8] - half[:, 0, 8].mean(); f = np.fft.rfftfreq(len(sig), dt); sp = np.abs(np.fft.rfft(sig))
        print(json.dumps(dict(cmd=target, falls=falls, vx_mean=float(half[:, :, 19].mean()), vx_std=float(half[:, :, 19].std()), z_mean=float(half[:, :, 2].mean()),
                              z_std=float(half[:, :, 2].std()), stride_hz=float(f[1:][np.argmax(sp[1:])]), reward_mean=float(rec["reward"][T // 2:].mean()),
                              max_joint_effort=float(np.abs(rec["effort"]).max()))))
    return 0


def main(argv=None):
    args = parse_args(argv if argv is not None else sys.argv[1:])
    if not args.train:
        return test_mode(args)
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
