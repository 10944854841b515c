"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement (NumPy float64) of the reference's PPO
update path.  The product package never imports this file.

Parity pin: TensorFlow / tensorflow-probability are not installable here and the reference ships
no logged PPO outputs, so PPO numerics are pinned by (a) ``compute_gae`` below being the
reference's own expression (scipy.signal.lfilter, utils.py:45-50) cross-checked against an
explicit backward recursion, (b) the shipped agent checkpoints loading into exactly these 13
tensor shapes, and (c) an independent torch-autograd float64 restatement (oracle/torch_ref.py)
agreeing with the hand-written backward to ~1e-12 (tests/test_oracle_ppo.py).  Status:
"parity unpinned against TF outputs" for PPO -- only self-consistency pins exist.

What it follows (reference repo root):
  * networks    ppo.py:38-66 + utils.py:25-28: pi 67->500->300 (relu, relu) -> dense A + tanh ->
                low + (t+1)/2 * (high-low); state-independent action_logstd[A]; SEPARATE value
                trunk 67->500->300 (relu, relu) -> 1
  * log-prob    tfp Normal.log_prob summed over actions, keepdims -> [B,1]
                = sum_a [-0.5 ((a-mu)/sigma)^2 - (0.9189385175704956 + log sigma)]
  * losses      ppo.py:119-134
  * optimiser   ppo.py:142-144  Adam(lr * lr_decay**episode) on the 13 policy/ tensors only
  * driver      train.py:171-207 (GAE, returns, advantage normalisation, theta_old <- theta,
                epochs x shuffled minibatches, short last minibatch allowed)
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict

import numpy as np
import scipy.signal

LOG_SQRT_2PI = 0.9189385175704956
ENTROPY_CONST = 1.4189385175704956

PPO_TENSORS = ["dense/kernel", "dense/bias", "dense_1/kernel", "dense_1/bias",
               "action_mean/kernel", "action_mean/bias", "action_logstd",
               "dense_2/kernel", "dense_2/bias", "dense_3/kernel", "dense_3/bias",
               "value/kernel", "value/bias"]


def param_shapes(state_dim=67, num_actions=2, pi_hidden=(500, 300), vf_hidden=(500, 300)):
    s = OrderedDict()
    s["dense/kernel"] = (state_dim, pi_hidden[0]);        s["dense/bias"] = (pi_hidden[0],)
    s["dense_1/kernel"] = (pi_hidden[0], pi_hidden[1]);   s["dense_1/bias"] = (pi_hidden[1],)
    s["action_mean/kernel"] = (pi_hidden[1], num_actions); s["action_mean/bias"] = (num_actions,)
    s["action_logstd"] = (num_actions,)
    s["dense_2/kernel"] = (state_dim, vf_hidden[0]);      s["dense_2/bias"] = (vf_hidden[0],)
    s["dense_3/kernel"] = (vf_hidden[0], vf_hidden[1]);   s["dense_3/bias"] = (vf_hidden[1],)
    s["value/kernel"] = (vf_hidden[1], 1);                s["value/bias"] = (1,)
    return s


def init_params(seed=0, state_dim=67, num_actions=2, initial_std=0.4, initial_mean_factor=0.1, dtype=np.float32):
    """tf.layers.dense default glorot-uniform; action_mean uses variance_scaling(scale=0.1)
    (fan_in, truncated normal: stddev = sqrt(scale/fan_in)/0.87962566103423978); logstd = log(initial_std)."""
    rng = np.random.RandomState(seed)
    out = OrderedDict()
    for name, shape in param_shapes(state_dim, num_actions).items():
        if name == "action_logstd":
            out[name] = np.full(shape, np.log(initial_std), dtype)
        elif name.endswith("bias"):
            out[name] = np.zeros(shape, dtype)
        elif name == "action_mean/kernel":
            std = np.sqrt(initial_mean_factor / shape[0]) / 0.87962566103423978
            t = rng.randn(*shape)
            bad = np.abs(t) > 2
            while bad.any():
                t[bad] = rng.randn(int(bad.sum()))
                bad = np.abs(t) > 2
            out[name] = (t * std).astype(dtype)
        else:
            limit = np.sqrt(6.0 / (shape[0] + shape[1]))
            out[name] = rng.uniform(-limit, limit, size=shape).astype(dtype)
    return out


# ----------------------------------------------------------------------------- GAE (utils.py:45-50)
def compute_gae(rewards, values, bootstrap_values, terminals, gamma, lam):
    """The reference's own expression: the terminal mask enters delta only; the IIR accumulation
    y[n] = x[n] + gamma*lam*y[n-1] over the reversed deltas is NOT reset at terminals.  float64."""
    rewards = np.array(rewards, dtype=np.float64)
    values = np.array(list(np.asarray(values, dtype=np.float64)) + [float(bootstrap_values)])
    terminals = np.array(terminals, dtype=np.float64)
    deltas = rewards + (1.0 - terminals) * gamma * values[1:] - values[:-1]
    return scipy.signal.lfilter([1], [1, -gamma * lam], deltas[::-1], axis=0)[::-1]


def compute_gae_loop(rewards, values, bootstrap_values, terminals, gamma, lam):
    """Same recurrence written as an explicit backward loop (cross-check of the lfilter form)."""
    t_len = len(rewards)
    v = np.append(np.asarray(values, np.float64), float(bootstrap_values))
    adv = np.zeros(t_len)
    acc = 0.0
    for t in range(t_len - 1, -1, -1):
        delta = float(rewards[t]) + (1.0 - float(terminals[t])) * gamma * v[t + 1] - v[t]
        acc = delta + gamma * lam * acc
        adv[t] = acc
    return adv


def returns_and_normalised_advantages(rewards, values, last_value, dones, gamma, lam):
    """train.py:175-177."""
    adv = compute_gae(rewards, values, last_value, dones, gamma, lam)
    returns = adv + np.asarray(values, np.float64)
    adv_n = (adv - adv.mean()) / (adv.std() + 1e-8)
    return returns, adv_n, adv


# ----------------------------------------------------------------------------- network
def forward(p: Dict[str, np.ndarray], s, low, high, keep=None):
    h1 = np.maximum(s @ p["dense/kernel"] + p["dense/bias"], 0.0)
    h2 = np.maximum(h1 @ p["dense_1/kernel"] + p["dense_1/bias"], 0.0)
    t = np.tanh(h2 @ p["action_mean/kernel"] + p["action_mean/bias"])
    mu = low + ((t + 1.0) / 2.0) * (high - low)
    g1 = np.maximum(s @ p["dense_2/kernel"] + p["dense_2/bias"], 0.0)
    g2 = np.maximum(g1 @ p["dense_3/kernel"] + p["dense_3/bias"], 0.0)
    v = (g2 @ p["value/kernel"] + p["value/bias"])[:, 0]
    if keep is not None:
        keep.update(h1=h1, h2=h2, t=t, g1=g1, g2=g2)
    return mu, v


def log_prob(mu, logstd, a):
    std = np.exp(logstd)
    return np.sum(-0.5 * ((a - mu) / std) ** 2 - (LOG_SQRT_2PI + logstd), axis=-1, keepdims=True)


def predict(p, s, low, high, noise=None):
    """PPO.predict (ppo.py:231-251): greedy when ``noise`` is None, else clip(mu + noise*sigma)."""
    s = np.asarray(s, np.float64)
    single = s.ndim != 2
    if single:
        s = s[None]
    mu, v = forward(p, s, low, high)
    act = mu if noise is None else np.clip(mu + np.asarray(noise) * np.exp(p["action_logstd"]), low, high)
    return (act[0], v[0]) if s.shape[0] == 1 else (act, v)


def loss_and_grads(params, params_old, s, a, ret, adv, low, high, epsilon=0.2, value_scale=0.5,
                   entropy_scale=0.01, want_grads=True, dtype=np.float64):
    """ppo.py:119-134 and its reverse-mode gradient w.r.t. the 13 policy tensors."""
    p = {k: np.asarray(v, dtype) for k, v in params.items()}
    po = {k: np.asarray(v, dtype) for k, v in params_old.items()}
    s = np.asarray(s, dtype); a = np.asarray(a, dtype); ret = np.asarray(ret, dtype); adv = np.asarray(adv, dtype)
    low = np.asarray(low, dtype); high = np.asarray(high, dtype)
    bsz = s.shape[0]
    # the graph's Const nodes are float32 (shipped .meta: clip_by_value/y = 0.800000011920929, clip_by_value/Minimum/y =
    # 1.2000000476837158, mul_3/y = 0.009999999776482582; tests/test_oracle.py pins them): use the rounded values
    clip_lo, clip_hi = float(np.float32(1.0 - epsilon)), float(np.float32(1.0 + epsilon))
    value_scale, entropy_scale = float(np.float32(value_scale)), float(np.float32(entropy_scale))
    keep = {}
    mu, v = forward(p, s, low, high, keep)
    mu_old, _ = forward(po, s, low, high)
    logstd = p["action_logstd"]
    std = np.exp(logstd)
    logp = log_prob(mu, logstd, a)
    logp_old = log_prob(mu_old, po["action_logstd"], a)
    ratio = np.exp(logp - logp_old)                       # [B,1]
    advc = adv[:, None]
    unclipped = ratio * advc
    clipped = np.clip(ratio, clip_lo, clip_hi) * advc
    policy_loss = np.mean(np.minimum(unclipped, clipped))
    value_loss = np.mean((v - ret) ** 2) * value_scale
    entropy_loss = np.sum(ENTROPY_CONST + logstd) * entropy_scale
    loss = -policy_loss + value_loss - entropy_loss
    out = dict(mu=mu, value=v, logp=logp, ratio=ratio, policy_loss=policy_loss, value_loss=value_loss,
               entropy_loss=entropy_loss, loss=loss, mean_ratio=ratio.mean())
    if not want_grads:
        return out

    g = {}
    # d(-policy_loss)/d ratio: tf.minimum routes to the first argument when unclipped <= clipped;
    # otherwise to the clipped branch whose own gradient is zero outside [1-eps, 1+eps]
    # (inside it, clipped == unclipped and the first branch already took it).
    first = unclipped <= clipped
    inside = (ratio >= clip_lo) & (ratio <= clip_hi)
    dratio = np.where(first, advc, np.where(inside, advc, 0.0)) * (-1.0 / bsz)
    dlogp = dratio * ratio                                # [B,1]
    diff = (a - mu) / std                                 # [B,A]
    dmu = dlogp * diff / std
    g["action_logstd"] = np.sum(dlogp * (diff * diff - 1.0), axis=0) - entropy_scale
    dt = dmu * 0.5 * (high - low)
    dpre = dt * (1.0 - keep["t"] ** 2)
    g["action_mean/kernel"] = keep["h2"].T @ dpre
    g["action_mean/bias"] = dpre.sum(axis=0)
    dh2 = (dpre @ p["action_mean/kernel"].T) * (keep["h2"] > 0)
    g["dense_1/kernel"] = keep["h1"].T @ dh2
    g["dense_1/bias"] = dh2.sum(axis=0)
    dh1 = (dh2 @ p["dense_1/kernel"].T) * (keep["h1"] > 0)
    g["dense/kernel"] = s.T @ dh1
    g["dense/bias"] = dh1.sum(axis=0)
    dv = (value_scale * 2.0 / bsz) * (v - ret)            # [B]
    g["value/kernel"] = keep["g2"].T @ dv[:, None]
    g["value/bias"] = np.array([dv.sum()])
    dg2 = (dv[:, None] @ p["value/kernel"].T) * (keep["g2"] > 0)
    g["dense_3/kernel"] = keep["g1"].T @ dg2
    g["dense_3/bias"] = dg2.sum(axis=0)
    dg1 = (dg2 @ p["dense_3/kernel"].T) * (keep["g1"] > 0)
    g["dense_2/kernel"] = s.T @ dg1
    g["dense_2/bias"] = dg1.sum(axis=0)
    out["grads"] = g
    return out


def learn(params, adam_state, states, actions, values, rewards, dones, last_value, low, high,
          gamma=0.99, lam=0.95, lr=1e-4, epsilon=0.2, value_scale=1.0, entropy_scale=0.01,
          num_epochs=3, batch_size=32, perms=None, dtype=np.float64):
    """train.py:171-207 with the minibatch permutations as an input (``perms[e]`` = index order of
    epoch e).  ``params`` (dict of ``dtype`` arrays) and ``adam_state`` are updated in place.
    Returns per-minibatch loss records."""
    from oracle.vae_oracle import adam_apply
    returns, adv_n, _ = returns_and_normalised_advantages(rewards, values, last_value, dones, gamma, lam)
    states = np.asarray(states, dtype); actions = np.asarray(actions, dtype)
    # the reference feeds float32 placeholders: returns/advantages are rounded on feed
    returns32 = returns.astype(np.float32).astype(dtype)
    adv32 = adv_n.astype(np.float32).astype(dtype)
    old = {k: v.copy() for k, v in params.items()}        # update_old_policy()
    n = states.shape[0]
    records = []
    for e in range(num_epochs):
        idx = np.asarray(perms[e])
        for i in range(int(np.ceil(n / batch_size))):
            mb = idx[i * batch_size:(i + 1) * batch_size]
            out = loss_and_grads(params, old, states[mb], actions[mb], returns32[mb], adv32[mb], low, high,
                                 epsilon, value_scale, entropy_scale, True, dtype)
            adam_apply(params, out["grads"], adam_state, lr)
            records.append((out["policy_loss"], out["value_loss"], out["entropy_loss"], out["loss"], out["mean_ratio"]))
    return records
