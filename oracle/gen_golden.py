#!/usr/bin/env python
"""Generates tests/golden/*.npz from the reference's own shipped artefacts and Python code.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):
    python oracle/gen_golden.py
Outputs (committed):
  tests/golden/trot_ref.npz          -- Exp_Raw_Data/trot_ref_.csv (q, dq, x; 10000 rows), the golden vector of
                                        gait_generator_manual + inverse_kinematics (Environment.hpp:1756-1890, 1687-1751)
  tests/golden/bp5_155_params.npz    -- the 19 parameter arrays of IRRL/script/pkl/bp5_155.pkl (both LSTM towers, heads,
                                        logstd) in tf.trainable_variables() order, plus the pi-tower CSV export
  tests/golden/lstm_kat.npz          -- inputs and outputs of the reference's own numpy policy
                                        (IRRL/script/utils/CustomerLstmNN.py:112-175 `predict`) run here on a 64-step
                                        input sequence with the CSV weights (pi tower) and with the pkl weights (pi + V)
"""
import io
import os
import pickle
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


class _Dummy:
    """Stands in for every non-numpy class/function referenced by the cloudpickle stream."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __setstate__(self, s):
        self.__dict__["_state"] = s

    def __getattr__(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Dummy()

    def __setitem__(self, k, v):
        pass

    def append(self, *a):
        pass

    def extend(self, *a):
        pass

    def update(self, *a, **k):
        pass


class _StubUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.split(".")[0] in ("numpy", "builtins", "collections", "_codecs", "copyreg") and name != "eval":
            try:
                return super().find_class(module, name)
            except Exception:
                pass
        return _Dummy


def load_pkl_params(path):
    with open(path, "rb") as f:
        data, params = _StubUnpickler(io.BytesIO(f.read())).load()
    params = [np.asarray(p, dtype=np.float32) for p in params]
    meta = {k: v for k, v in data.items() if isinstance(v, (int, float))}
    return meta, params


def main():
    os.makedirs(OUT, exist_ok=True)
    # ---- 1. gait golden vector
    raw = np.loadtxt(os.path.join(REF, "Exp_Raw_Data", "trot_ref_.csv"), skiprows=1)
    assert raw.shape == (10000, 28), raw.shape
    np.savez_compressed(os.path.join(OUT, "trot_ref.npz"), x=raw[:, 0].astype(np.float32), z=raw[:, 1].astype(np.float32),
                        q=raw[:, 3:15].astype(np.float32), dq=raw[:, 15:27].astype(np.float32))
    # ---- 2. weights
    meta, params = load_pkl_params(os.path.join(REF, "IRRL", "script", "pkl", "bp5_155.pkl"))
    names = ["lstm_pi0_wx", "lstm_pi0_wh", "lstm_pi0_b", "lstm_pi1_wx", "lstm_pi1_wh", "lstm_pi1_b",
             "lstm_v0_wx", "lstm_v0_wh", "lstm_v0_b", "lstm_v1_wx", "lstm_v1_wh", "lstm_v1_b",
             "vf_w", "vf_b", "pi_w", "pi_b", "pi_logstd", "q_w", "q_b"]
    shapes = [p.shape for p in params]
    print("pkl meta", meta)
    print("pkl shapes", shapes)
    assert len(params) == 19
    csvdir = os.path.join(REF, "IRRL", "script", "model", "bp5_155")
    csv = {n: np.loadtxt(os.path.join(csvdir, n + ".csv"), delimiter=",").astype(np.float32)
           for n in ("lstm_wx0", "lstm_wh0", "lstm_b0", "lstm_wx1", "lstm_wh1", "lstm_b1", "pi_w", "pi_b")}
    np.savez_compressed(os.path.join(OUT, "bp5_155_params.npz"), **{n: p for n, p in zip(names, params)},
                        **{"csv_" + k: v for k, v in csv.items()},
                        meta_keys=np.array(list(meta.keys())), meta_vals=np.array([float(v) for v in meta.values()]))
    # ---- 3. LSTM known answers from the reference's own numpy class
    sys.modules["raisim_gym"] = types.ModuleType("raisim_gym")
    sys.modules["raisim_gym.algo"] = types.ModuleType("raisim_gym.algo")
    m = types.ModuleType("raisim_gym.algo.ppo2")
    m.PPO2 = _Dummy
    sys.modules["raisim_gym.algo.ppo2"] = m
    sys.path.insert(0, os.path.join(REF, "IRRL", "script"))
    cwd = os.getcwd()
    os.chdir(os.path.join(REF, "IRRL", "script"))   # the class resolves os.getcwd() + '/model/<name>' (CustomerLstmNN.py:34-36)
    try:
        from utils.CustomerLstmNN import CustomerLstmNN
        pol = CustomerLstmNN("./pkl/bp5_155.pkl", n_lstm=[48, 48])
    finally:
        os.chdir(cwd)
    assert len(pol.lstm_wx) == 2 and pol.lstm_wx[0].shape == (35, 192)
    # the 35-vector of CustomerLstmNN.py:234-244 (its own demo input)
    kat_in = np.array([-0.25, 0, 0, -1, 1, 0.0495351, -0.0115055, 0.10003, 0.0494637, -0.0106432, 0.0999373, -0.0649202,
                       -0.0107303, 0.110537, 0.064962, -0.010784, 0.110514, -0.00990702, -0.158301, 0.334006, 0.00989274,
                       -0.158129, 0.333987, -0.012984, -0.158146, 0.336107, 0.0129924, -0.158157, 0.336103, 0.0144317,
                       0.00293443, -7.58682e-05, 0.00637404, 0.000356918, -0.00200364])
    rng = np.random.default_rng(155)
    seq = np.concatenate([kat_in[None], rng.normal(0, 0.5, size=(63, 35))], 0)
    outs, hid = [], []
    pol.reset()
    for x in seq:
        outs.append(pol.predict(x))
        hid.append(pol.get_hidden_state())
    # second pass: same class, pkl weights for both towers (flag_v path, CustomerLstmNN.py:137-156)
    pol2 = CustomerLstmNN.__new__(CustomerLstmNN)
    pol2.n_lstm = [48, 48]; pol2.flag_v = True; pol2.is_LSTM = True
    P = dict(zip(names, params))
    pol2.lstm_wx = [P["lstm_pi0_wx"].astype(np.float64), P["lstm_pi1_wx"].astype(np.float64)]
    pol2.lstm_wh = [P["lstm_pi0_wh"].astype(np.float64), P["lstm_pi1_wh"].astype(np.float64)]
    pol2.lstm_b = [P["lstm_pi0_b"].astype(np.float64), P["lstm_pi1_b"].astype(np.float64)]
    pol2.v_lstm_wx = [P["lstm_v0_wx"].astype(np.float64), P["lstm_v1_wx"].astype(np.float64)]
    pol2.v_lstm_wh = [P["lstm_v0_wh"].astype(np.float64), P["lstm_v1_wh"].astype(np.float64)]
    pol2.v_lstm_b = [P["lstm_v0_b"].astype(np.float64), P["lstm_v1_b"].astype(np.float64)]
    pol2.pi_w, pol2.pi_b = P["pi_w"].astype(np.float64), P["pi_b"].astype(np.float64)
    pol2.v_w, pol2.v_b = P["vf_w"].astype(np.float64), P["vf_b"].astype(np.float64)
    pol2.cell_state = [np.zeros(48), np.zeros(48)]; pol2.hidden = [np.zeros(48), np.zeros(48)]
    pol2.v_cell_state = [np.zeros(48), np.zeros(48)]; pol2.v_hidden = [np.zeros(48), np.zeros(48)]
    outs2, vals2, hid2, vhid2 = [], [], [], []
    for x in seq:
        outs2.append(pol2.predict(x)); vals2.append(np.array(pol2.get_v()).reshape(-1)[0]); hid2.append(pol2.get_hidden_state())
        vhid2.append(np.hstack((pol2.v_cell_state[0], pol2.v_hidden[0], pol2.v_cell_state[1], pol2.v_hidden[1])))
    np.savez_compressed(os.path.join(OUT, "lstm_kat.npz"), seq=seq, csv_action=np.array(outs), csv_state=np.array(hid),
                        pkl_action=np.array(outs2), pkl_value=np.array(vals2), pkl_state_pi=np.array(hid2), pkl_state_v=np.array(vhid2))
    print("first clipped pi mean (CSV weights):", np.round(outs[0], 5))
    print("max |csv - pkl| action:", np.abs(np.array(outs) - np.array(outs2)).max())
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
