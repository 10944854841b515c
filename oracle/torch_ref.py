"""ORACLE — TEST INFRASTRUCTURE ONLY.  torch-CPU restatement of the reference graphs built from
stock torch ops + autograd.  Two uses:

  1. float64: an INDEPENDENT derivation of the gradients (autograd instead of the hand-written
     backward in vae_oracle.py / ppo_oracle.py); tests require both to agree to ~1e-12.
  2. float32 with all host threads (oneDNN): the timed CPU baseline of bench.py
     (``cpu_baseline`` / ``--impl reference``).  TensorFlow 1.13 cannot be installed here, so this
     is the closest runnable stand-in for "the reference's own CPU path" (kind = "port").

Follows reference vae/models.py:85-142, 249-266 and ppo.py:38-66, 119-147 (see vae_oracle.py /
ppo_oracle.py for the line-by-line notes).  The product package never imports this file.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F


def _t(a, dtype):
    return torch.as_tensor(np.asarray(a), dtype=dtype)


def vae_params_to_torch(params: Dict[str, np.ndarray], dtype=torch.float64, requires_grad=True):
    out = {}
    for k, v in params.items():
        t = _t(v, dtype).clone()
        t.requires_grad_(requires_grad)
        out[k] = t
    return out


def vae_forward(p, x, eps, training=True):
    """x [B,H,W,3] (NHWC, like the reference placeholders) -> mean, logvar, z, logits[B,H,W,C_t]."""
    a = x.permute(0, 3, 1, 2)
    for name in ("conv1", "conv2", "conv3", "conv4"):
        w = p["encoder/%s/kernel" % name].permute(3, 2, 0, 1)          # [kh,kw,ci,co] -> [co,ci,kh,kw]
        a = F.relu(F.conv2d(a, w, p["encoder/%s/bias" % name], stride=2))
    eh, ew = a.shape[2], a.shape[3]
    flat = a.permute(0, 2, 3, 1).reshape(a.shape[0], -1)               # NHWC flatten order
    mean = flat @ p["mean/kernel"] + p["mean/bias"]
    logvar = flat @ p["logstd_sqare/kernel"] + p["logstd_sqare/bias"]
    z = mean + eps * torch.exp(0.5 * logvar) if training else mean
    d = z @ p["decoder/dense1/kernel"] + p["decoder/dense1/bias"]
    a = d.reshape(-1, eh, ew, 256).permute(0, 3, 1, 2)
    for name in ("deconv1", "deconv2", "deconv3", "deconv4"):
        w = p["decoder/%s/kernel" % name].permute(3, 2, 0, 1)          # [kh,kw,co,ci] -> [ci,co,kh,kw]
        a = F.conv_transpose2d(a, w, p["decoder/%s/bias" % name], stride=2)
        if name != "deconv4":
            a = F.relu(a)
    logits = a.permute(0, 2, 3, 1)
    return mean, logvar, z, logits


def vae_loss(p, x, y, eps, loss_type="mse", beta=1.0, kl_tolerance=0.0):
    mean, logvar, z, logits = vae_forward(p, x, eps)
    b = x.shape[0]
    lf = logits.reshape(b, -1)
    yf = y.reshape(b, -1)
    if loss_type == "mse":
        elem = (yf - torch.sigmoid(lf)) ** 2
    elif loss_type == "bce":
        elem = F.binary_cross_entropy_with_logits(lf, yf, reduction="none")
    elif loss_type == "bce_v2":
        s = torch.sigmoid(lf)
        elem = -(yf * torch.log(1e-10 + s) + (1 - yf) * torch.log(1e-10 + 1 - s))
    else:
        raise ValueError(loss_type)
    recon = elem.sum(dim=1).mean()
    kl_rows = -0.5 * torch.sum(1.0 + logvar - mean * mean - torch.exp(logvar), dim=1)
    if kl_tolerance > 0:
        kl_rows = torch.maximum(kl_rows, torch.full_like(kl_rows, kl_tolerance * mean.shape[1]))
    kl = kl_rows.mean()
    return recon + beta * kl, recon, kl, (mean, logvar, z, logits)


def vae_loss_and_grads(params, x, y, eps, loss_type="mse", beta=1.0, kl_tolerance=0.0, dtype=torch.float64):
    p = vae_params_to_torch(params, dtype)
    loss, recon, kl, (mean, logvar, z, logits) = vae_loss(p, _t(x, dtype), _t(y, dtype), _t(eps, dtype),
                                                          loss_type, beta, kl_tolerance)
    loss.backward()
    return dict(mean=mean.detach().numpy(), logvar=logvar.detach().numpy(), z=z.detach().numpy(),
                logits=logits.detach().numpy(), recon=float(recon.detach()), kl=float(kl.detach()), loss=float(loss.detach()),
                grads={k: v.grad.numpy() for k, v in p.items()})


class TorchAdamTF:
    """TF ApplyAdam on a dict of torch tensors (same maths as vae_oracle.adam_apply)."""

    def __init__(self, p, lr, beta1=0.9, beta2=0.999, eps=1e-8):
        self.p = p
        # float32-rounded graph constants, like vae_oracle.ADAM_* (pinned to the shipped .meta files)
        f32 = lambda v: float(np.float32(v))
        self.lr, self.b1, self.b2, self.eps = f32(lr), f32(beta1), f32(beta2), f32(eps)
        self.m = {k: torch.zeros_like(v) for k, v in p.items()}
        self.v = {k: torch.zeros_like(v) for k, v in p.items()}
        self.b1p, self.b2p = self.b1, self.b2

    @torch.no_grad()
    def step(self, lr=None):
        lr = self.lr if lr is None else float(np.float32(lr))
        alpha = lr * math.sqrt(1.0 - self.b2p) / (1.0 - self.b1p)
        for k, t in self.p.items():
            if t.grad is None:
                continue
            g = t.grad
            self.m[k].add_((g - self.m[k]) * (1.0 - self.b1))
            self.v[k].add_((g * g - self.v[k]) * (1.0 - self.b2))
            t.sub_(alpha * self.m[k] / (self.v[k].sqrt() + self.eps))
            t.grad = None
        self.b1p *= self.b1
        self.b2p *= self.b2


class TorchVAETrainer:
    """fp32 CPU train step used as the timed baseline: forward + loss + backward + TF-Adam."""

    def __init__(self, params, lr=1e-4, loss_type="mse", beta=1.0, kl_tolerance=0.0, dtype=torch.float32):
        self.dtype = dtype
        self.p = vae_params_to_torch(params, dtype)
        self.opt = TorchAdamTF(self.p, lr)
        self.loss_type, self.beta, self.kl_tolerance = loss_type, beta, kl_tolerance

    def step(self, x, y, eps, micro_batch=None):
        """One optimiser step on the whole batch; ``micro_batch`` splits it into gradient-accumulation
        chunks (identical result up to fp32 summation order) so a 4096-frame batch fits any host."""
        b = x.shape[0]
        mb = b if micro_batch is None else micro_batch
        recon_sum = kl_sum = 0.0
        for s in range(0, b, mb):
            xs, ys, es = x[s:s + mb], y[s:s + mb], eps[s:s + mb]
            loss, recon, kl, _ = vae_loss(self.p, xs, ys, es, self.loss_type, self.beta, self.kl_tolerance)
            (loss * (xs.shape[0] / b)).backward()
            recon_sum += float(recon) * xs.shape[0]
            kl_sum += float(kl) * xs.shape[0]
        self.opt.step()
        return recon_sum / b, kl_sum / b


# ----------------------------------------------------------------------------- PPO
_LOG_SQRT_2PI = 0.9189385175704956
_ENTROPY_CONST = 1.4189385175704956


def ppo_params_to_torch(params, dtype=torch.float64, requires_grad=True):
    return vae_params_to_torch(params, dtype, requires_grad)


def ppo_forward(p, s, low, high):
    """ppo.py:38-66.  Returns action_mean [B,A], value [B]."""
    h = F.relu(s @ p["dense/kernel"] + p["dense/bias"])
    h = F.relu(h @ p["dense_1/kernel"] + p["dense_1/bias"])
    t = torch.tanh(h @ p["action_mean/kernel"] + p["action_mean/bias"])
    mu = low + ((t + 1) / 2) * (high - low)
    hv = F.relu(s @ p["dense_2/kernel"] + p["dense_2/bias"])
    hv = F.relu(hv @ p["dense_3/kernel"] + p["dense_3/bias"])
    v = (hv @ p["value/kernel"] + p["value/bias"]).squeeze(-1)
    return mu, v


def ppo_logp(mu, logstd, a):
    std = torch.exp(logstd)
    return torch.sum(-0.5 * ((a - mu) / std) ** 2 - (_LOG_SQRT_2PI + logstd), dim=-1, keepdim=True)


def ppo_loss(p, p_old, s, a, ret, adv, low, high, epsilon=0.2, value_scale=0.5, entropy_scale=0.01):
    """ppo.py:119-134.  Returns loss, (policy_loss, value_loss, entropy_loss, mean_ratio)."""
    # float32-rounded graph constants (see ppo_oracle.loss_and_grads; pinned to the shipped .meta)
    clip_lo, clip_hi = float(np.float32(1.0 - epsilon)), float(np.float32(1.0 + epsilon))
    value_scale, entropy_scale = float(np.float32(value_scale)), float(np.float32(entropy_scale))
    mu, v = ppo_forward(p, s, low, high)
    with torch.no_grad():
        mu_old, _ = ppo_forward(p_old, s, low, high)
        logp_old = ppo_logp(mu_old, p_old["action_logstd"], a)
    logp = ppo_logp(mu, p["action_logstd"], a)
    ratio = torch.exp(logp - logp_old)
    advc = adv.unsqueeze(-1)
    policy_loss = torch.mean(torch.minimum(ratio * advc, torch.clamp(ratio, clip_lo, clip_hi) * advc))
    value_loss = torch.mean((v - ret) ** 2) * value_scale
    entropy = torch.sum(_ENTROPY_CONST + p["action_logstd"]).expand(s.shape[0])
    entropy_loss = torch.mean(entropy) * entropy_scale
    loss = -policy_loss + value_loss - entropy_loss
    return loss, (policy_loss, value_loss, entropy_loss, ratio.mean())


def ppo_loss_and_grads(params, params_old, s, a, ret, adv, low, high, epsilon=0.2, value_scale=0.5,
                       entropy_scale=0.01, dtype=torch.float64):
    p = ppo_params_to_torch(params, dtype)
    po = ppo_params_to_torch(params_old, dtype, requires_grad=False)
    loss, (pl, vl, el, mr) = ppo_loss(p, po, _t(s, dtype), _t(a, dtype), _t(ret, dtype), _t(adv, dtype),
                                      _t(low, dtype), _t(high, dtype), epsilon, value_scale, entropy_scale)
    loss.backward()
    return dict(loss=float(loss.detach()), policy_loss=float(pl.detach()), value_loss=float(vl.detach()), entropy_loss=float(el.detach()),
                mean_ratio=float(mr.detach()), grads={k: v.grad.numpy() for k, v in p.items()})


class TorchPPOLearner:
    """fp32 CPU restatement of the driver's update block (train.py:171-207), timed baseline."""

    def __init__(self, params, low, high, lr=1e-4, epsilon=0.2, value_scale=1.0, entropy_scale=0.01,
                 dtype=torch.float32):
        self.dtype = dtype
        self.p = ppo_params_to_torch(params, dtype)
        self.p_old = ppo_params_to_torch(params, dtype, requires_grad=False)
        self.low, self.high = _t(low, dtype), _t(high, dtype)
        self.opt = TorchAdamTF(self.p, lr)
        self.epsilon, self.value_scale, self.entropy_scale = epsilon, value_scale, entropy_scale

    def learn(self, states, actions, values, rewards, dones, last_value, gamma, lam, num_epochs, batch_size, perms):
        from oracle.ppo_oracle import compute_gae
        adv = compute_gae(rewards, values, last_value, dones, gamma, lam)
        ret = adv + np.asarray(values)
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)
        s, a = _t(states, self.dtype), _t(actions, self.dtype)
        r, ad = _t(ret, self.dtype), _t(adv, self.dtype)
        with torch.no_grad():
            for k in self.p:
                self.p_old[k].copy_(self.p[k])
        n = s.shape[0]
        for e in range(num_epochs):
            idx = torch.as_tensor(np.asarray(perms[e]))
            for i in range(int(np.ceil(n / batch_size))):
                mb = idx[i * batch_size:(i + 1) * batch_size]
                loss, _ = ppo_loss(self.p, self.p_old, s[mb], a[mb], r[mb], ad[mb], self.low, self.high,
                                   self.epsilon, self.value_scale, self.entropy_scale)
                loss.backward()
                self.opt.step()


# ----------------------------------------------------------------------------- MlpVAE (independent autograd derivation)
def mlp_vae_loss_and_grads(params, x, y, eps, loss_type="mse", beta=1.0, kl_tolerance=0.0, dtype=torch.float64):
    """vae/models.py:271-299 with stock torch ops + autograd; the check of vae_oracle.mlp_loss_and_grads' hand-written backward."""
    p = vae_params_to_torch(params, dtype)
    xt, yt, et = _t(x, dtype), _t(y, dtype), _t(eps, dtype)
    b = xt.shape[0]
    h = F.relu(xt.reshape(b, -1) @ p["encoder/dense/kernel"] + p["encoder/dense/bias"])
    h = F.relu(h @ p["encoder/dense_1/kernel"] + p["encoder/dense_1/bias"])
    mean = h @ p["mean/kernel"] + p["mean/bias"]
    logvar = h @ p["logstd_sqare/kernel"] + p["logstd_sqare/bias"]
    z = mean + et * torch.exp(0.5 * logvar)
    g = F.relu(z @ p["decoder/dense/kernel"] + p["decoder/dense/bias"])
    g = F.relu(g @ p["decoder/dense_1/kernel"] + p["decoder/dense_1/bias"])
    lf = g @ p["decoder/dense_2/kernel"] + p["decoder/dense_2/bias"]
    yf = yt.reshape(b, -1)
    if loss_type == "mse":
        elem = (yf - torch.sigmoid(lf)) ** 2
    elif loss_type == "bce":
        elem = F.binary_cross_entropy_with_logits(lf, yf, reduction="none")
    else:
        sg = torch.sigmoid(lf)
        elem = -(yf * torch.log(1e-10 + sg) + (1 - yf) * torch.log(1e-10 + 1 - sg))
    recon = elem.sum(dim=1).mean()
    kl_rows = -0.5 * torch.sum(1.0 + logvar - mean * mean - torch.exp(logvar), dim=1)
    if kl_tolerance > 0:
        kl_rows = torch.maximum(kl_rows, torch.full_like(kl_rows, kl_tolerance * mean.shape[1]))
    kl = kl_rows.mean()
    (recon + beta * kl).backward()
    return dict(mean=mean.detach().numpy(), logvar=logvar.detach().numpy(), logits=lf.detach().numpy(), recon=float(recon.detach()),
                kl=float(kl.detach()), grads={k: v.grad.numpy() for k, v in p.items()})
