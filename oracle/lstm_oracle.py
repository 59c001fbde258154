"""CPU ORACLE (test infrastructure only) for the two-tower 2x48 LSTM policy act().

numpy restatement of
  * IRRL/script/utils/CustomerLstmNN.py:112-156 (`predict`: gate order i,f,o,g; c = s(f) c + s(i) tanh(g);
    h = s(o) tanh(c); pi head; V tower with the same cell),
  * IRRL/script/run_bp_v5.py:117-176 (`CustomLSTMPolicy`: tower wiring, state layout
    [c0,h0,c1,h1]_pi || [c0,h0,c1,h1]_V = 384 floats, masks zero the state where done(t-1)),
  * stable-baselines 2.8.0 (third party, absent from /root/reference; semantics per SURVEY.md 9.8):
    `lstm()` multiplies c and h by (1 - mask) before the cell; `DiagGaussianProbabilityDistribution`:
    a = mu + exp(logstd) * eps, neglogp = 0.5 sum(((a-mu)/sigma)^2) + 0.5*12*ln(2 pi) + sum(logstd).
Pinned by tests/golden/lstm_kat.npz (outputs of the reference's own CustomerLstmNN run on bp5_155 weights).
"""
import numpy as np

N_LSTM = 48
STATE_DIM = 8 * N_LSTM   # 384
PARAM_NAMES = ["lstm_pi0_wx", "lstm_pi0_wh", "lstm_pi0_b", "lstm_pi1_wx", "lstm_pi1_wh", "lstm_pi1_b",
               "lstm_v0_wx", "lstm_v0_wh", "lstm_v0_b", "lstm_v1_wx", "lstm_v1_wh", "lstm_v1_b",
               "vf_w", "vf_b", "pi_w", "pi_b", "pi_logstd", "q_w", "q_b"]


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def lstm_cell(x, c, h, wx, wh, b):
    """CustomerLstmNN.py:116-128 for a batch: x [N,in], c/h [N,48]."""
    gate = x @ wx + h @ wh + b
    n = c.shape[1]
    i = _sigmoid(gate[:, 0:n]); f = _sigmoid(gate[:, n:2 * n]); o = _sigmoid(gate[:, 2 * n:3 * n]); g = np.tanh(gate[:, 3 * n:4 * n])
    c2 = f * c + i * g
    h2 = o * np.tanh(c2)
    return c2, h2


def act(params, obs, state, mask, eps=None, dtype=np.float64):
    """One policy step (run_bp_v5.py:178-185).

    params: dict name->array; obs [N,35]; state [N,384]; mask [N] (1.0 where the previous step was done);
    eps [N,12] standard normal draws (None -> deterministic action = mean).
    Returns action [N,12] (unclipped), value [N], new_state [N,384], neglogp [N], mean [N,12].
    """
    P = {k: np.asarray(v, dtype) for k, v in params.items()}
    obs = np.asarray(obs, dtype); state = np.asarray(state, dtype).copy()
    keep = (1.0 - np.asarray(mask, dtype))[:, None]
    state = state * keep                                  # SB lstm(): c *= 1-m ; h *= 1-m
    n = N_LSTM
    new_state = np.zeros_like(state)
    lat = obs
    for li, pre in enumerate(("lstm_pi0", "lstm_pi1")):
        c = state[:, (2 * li) * n:(2 * li + 1) * n]; h = state[:, (2 * li + 1) * n:(2 * li + 2) * n]   # [c,h] order (CustomerLstmNN.py:189)
        c, h = lstm_cell(lat, c, h, P[pre + "_wx"], P[pre + "_wh"], P[pre + "_b"])
        new_state[:, (2 * li) * n:(2 * li + 1) * n] = c; new_state[:, (2 * li + 1) * n:(2 * li + 2) * n] = h
        lat = h
    mean = lat @ P["pi_w"] + P["pi_b"]
    latv = obs
    for li, pre in enumerate(("lstm_v0", "lstm_v1")):
        o = 4 * n
        c = state[:, o + (2 * li) * n:o + (2 * li + 1) * n]; h = state[:, o + (2 * li + 1) * n:o + (2 * li + 2) * n]
        c, h = lstm_cell(latv, c, h, P[pre + "_wx"], P[pre + "_wh"], P[pre + "_b"])
        new_state[:, o + (2 * li) * n:o + (2 * li + 1) * n] = c; new_state[:, o + (2 * li + 1) * n:o + (2 * li + 2) * n] = h
        latv = h
    value = (latv @ P["vf_w"] + P["vf_b"])[:, 0]
    logstd = P["pi_logstd"].reshape(1, -1)
    std = np.exp(logstd)
    action = mean if eps is None else mean + std * np.asarray(eps, dtype)
    neglogp = 0.5 * np.sum(((action - mean) / std) ** 2, axis=1) + 0.5 * np.log(2.0 * np.pi) * mean.shape[1] + np.sum(logstd, axis=1)
    return action, value, new_state, neglogp, mean


def gae(rewards, values, dones, last_values, last_dones, gamma=0.99, lam=0.998):
    """ppo2.py:554-568: rewards/values/dones [T,N]; dones[t] = done flag *before* step t (mb_dones)."""
    T = rewards.shape[0]
    adv = np.zeros_like(rewards)
    last = 0.0
    for t in reversed(range(T)):
        if t == T - 1:
            nonterm = 1.0 - last_dones; nextv = last_values
        else:
            nonterm = 1.0 - dones[t + 1]; nextv = values[t + 1]
        delta = rewards[t] + gamma * nextv * nonterm - values[t]
        adv[t] = last = delta + gamma * lam * nonterm * last
    return adv, adv + values
