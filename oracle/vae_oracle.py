"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement (NumPy, float64 by default) of the
reference's ConvVAE training graph.  Nothing in the product package may import this file;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg use it, and
only as the checker.

Parity pin: TensorFlow 1.13.1 is not installable in the build container, so this restatement
is pinned by the reference's shipped artefacts instead (tests/test_oracle_kat.py): shipped
checkpoints + shipped frames reproduce the shipped event-file losses (KAT-1/KAT-2), the
shipped Adam beta-powers match the step count, and an independent torch-autograd float64
restatement (oracle/torch_ref.py) agrees with the hand-derived backward below to ~1e-12.

What it follows (all paths relative to the reference repo root):
  * layers            vae/models.py:249-256 (encoder), :258-266 (decoder); TF-1.13 semantics
                      NHWC, VALID, stride 2, conv kernel [kh,kw,Cin,Cout], conv2d_transpose
                      kernel [kh,kw,Cout,Cin] (= Conv2DBackpropInput), dense kernel [in,out]
  * heads / sampling  vae/models.py:97-105   z = mean + eps * exp(0.5 * logstd_sq)
  * logits / sigmoid  vae/models.py:112-113
  * losses            vae/models.py:7-9 (KL), :11-22 (bce, bce_v2, mse), :122-137
  * optimiser         vae/models.py:140-142  tf.train.AdamOptimizer(constant lr).minimize
  * encode/decode     vae/models.py:188-202

Every weight tensor is kept in the TF layout; in the notation of this repo a stride-2 layer
connects a "big" image [Hb,Wb,Cb] and a "small" image [Hs,Ws,Cs] through a kernel
[k,k,Cb,Cs] -- that is the native layout of BOTH tf.layers.conv2d (big=input) and
tf.layers.conv2d_transpose (big=output).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Optional, Tuple

import numpy as np

Z_DIM_DEFAULT = 64

# (name, kind, kernel_size, Cb, Cs)  in TF variable-creation order
CONV_LAYERS = [
    ("encoder/conv1", 4, 3, 32),
    ("encoder/conv2", 4, 32, 64),
    ("encoder/conv3", 4, 64, 128),
    ("encoder/conv4", 4, 128, 256),
]


def param_shapes(source_shape=(80, 160, 3), target_channels=3, z_dim=Z_DIM_DEFAULT) -> "OrderedDict[str, Tuple[int, ...]]":
    """Variable names (without the leading ``vae/`` scope) and shapes, in creation order."""
    h, w, c = source_shape
    eh, ew = h, w
    for _ in range(4):
        eh, ew = (eh - 4) // 2 + 1, (ew - 4) // 2 + 1
    feat = eh * ew * 256
    s = OrderedDict()
    s["encoder/conv1/kernel"] = (4, 4, c, 32);     s["encoder/conv1/bias"] = (32,)
    s["encoder/conv2/kernel"] = (4, 4, 32, 64);    s["encoder/conv2/bias"] = (64,)
    s["encoder/conv3/kernel"] = (4, 4, 64, 128);   s["encoder/conv3/bias"] = (128,)
    s["encoder/conv4/kernel"] = (4, 4, 128, 256);  s["encoder/conv4/bias"] = (256,)
    s["mean/kernel"] = (feat, z_dim);              s["mean/bias"] = (z_dim,)
    s["logstd_sqare/kernel"] = (feat, z_dim);      s["logstd_sqare/bias"] = (z_dim,)
    s["decoder/dense1/kernel"] = (z_dim, feat);    s["decoder/dense1/bias"] = (feat,)
    s["decoder/deconv1/kernel"] = (4, 4, 128, 256); s["decoder/deconv1/bias"] = (128,)
    s["decoder/deconv2/kernel"] = (4, 4, 64, 128);  s["decoder/deconv2/bias"] = (64,)
    s["decoder/deconv3/kernel"] = (5, 5, 32, 64);   s["decoder/deconv3/bias"] = (32,)
    s["decoder/deconv4/kernel"] = (4, 4, target_channels, 32); s["decoder/deconv4/bias"] = (target_channels,)
    return s


def glorot_init(seed=0, source_shape=(80, 160, 3), target_channels=3, z_dim=Z_DIM_DEFAULT, dtype=np.float32):
    """tf.layers default init: glorot-uniform kernels, zero biases (reference never overrides)."""
    rng = np.random.RandomState(seed)
    out = OrderedDict()
    for name, shape in param_shapes(source_shape, target_channels, z_dim).items():
        if name.endswith("bias"):
            out[name] = np.zeros(shape, dtype)
            continue
        if len(shape) == 4:
            rf = shape[0] * shape[1]
            if "deconv" in name:       # conv2d_transpose kernel [kh,kw,out,in]
                fan_in, fan_out = shape[3] * rf, shape[2] * rf
            else:
                fan_in, fan_out = shape[2] * rf, shape[3] * rf
        else:
            fan_in, fan_out = shape
        limit = np.sqrt(6.0 / (fan_in + fan_out))
        out[name] = rng.uniform(-limit, limit, size=shape).astype(dtype)
    return out


# ----------------------------------------------------------------------------- primitives
def _windows(big: np.ndarray, k: int, hs: int, ws: int) -> np.ndarray:
    """[B,Hb,Wb,Cb] -> view [B,hs,ws,k,k,Cb] of the stride-2 windows big[b, 2i+kh, 2j+kw, c]."""
    b, hb, wb, cb = big.shape
    sb, sh, sw, sc = big.strides
    assert 2 * (hs - 1) + k <= hb and 2 * (ws - 1) + k <= wb
    return np.lib.stride_tricks.as_strided(
        big, shape=(b, hs, ws, k, k, cb), strides=(sb, 2 * sh, 2 * sw, sh, sw, sc), writeable=False)


def conv_gather(big: np.ndarray, w: np.ndarray) -> np.ndarray:
    """small[b,i,j,cs] = sum_{kh,kw,cb} big[b,2i+kh,2j+kw,cb] * w[kh,kw,cb,cs]
    (= tf Conv2D VALID stride 2; also the data-gradient of conv2d_transpose)."""
    k = w.shape[0]
    b, hb, wb, cb = big.shape
    hs, ws = (hb - k) // 2 + 1, (wb - k) // 2 + 1
    cols = _windows(big, k, hs, ws).reshape(b * hs * ws, k * k * cb)
    return (cols @ w.reshape(k * k * cb, -1)).reshape(b, hs, ws, -1)


def conv_scatter(small: np.ndarray, w: np.ndarray, out_hw: Optional[Tuple[int, int]] = None) -> np.ndarray:
    """big[b,2i+kh,2j+kw,cb] += small[b,i,j,cs] * w[kh,kw,cb,cs]
    (= tf Conv2DBackpropInput: conv2d_transpose forward, and the data-gradient of Conv2D)."""
    k = w.shape[0]
    b, hs, ws, cs = small.shape
    hb, wb = (hs - 1) * 2 + k, (ws - 1) * 2 + k
    if out_hw is not None:
        assert out_hw[0] >= hb and out_hw[1] >= wb
        hb_out, wb_out = out_hw
    else:
        hb_out, wb_out = hb, wb
    big = np.zeros((b, hb_out, wb_out, w.shape[2]), dtype=small.dtype)
    flat = small.reshape(-1, cs)
    for kh in range(k):
        for kw in range(k):
            contrib = (flat @ w[kh, kw].T).reshape(b, hs, ws, -1)
            big[:, kh:kh + 2 * hs:2, kw:kw + 2 * ws:2, :] += contrib
    return big


def conv_wgrad(big: np.ndarray, small: np.ndarray, k: int) -> np.ndarray:
    """gw[kh,kw,cb,cs] = sum_{b,i,j} big[b,2i+kh,2j+kw,cb] * small[b,i,j,cs]
    (= tf Conv2DBackpropFilter for both layer kinds)."""
    b, hs, ws, cs = small.shape
    cb = big.shape[3]
    cols = _windows(big, k, hs, ws).reshape(b * hs * ws, k * k * cb)
    return (cols.T @ small.reshape(-1, cs)).reshape(k, k, cb, cs)


def sigmoid(x):
    return 0.5 * (1.0 + np.tanh(0.5 * x))


# ----------------------------------------------------------------------------- forward
def encode(params: Dict[str, np.ndarray], x: np.ndarray, keep=None):
    """vae/models.py:249-256 + :97-98.  Returns (mean, logstd_sq)."""
    a = x
    for name in ("conv1", "conv2", "conv3", "conv4"):
        pre = conv_gather(a, params["encoder/%s/kernel" % name]) + params["encoder/%s/bias" % name]
        a = np.maximum(pre, 0.0)
        if keep is not None:
            keep[name] = a
            keep["pre/" + name] = pre
    flat = a.reshape(a.shape[0], -1)            # tf.layers.flatten of NHWC -> (h, w, c) order
    mean = flat @ params["mean/kernel"] + params["mean/bias"]
    logvar = flat @ params["logstd_sqare/kernel"] + params["logstd_sqare/bias"]
    return mean, logvar


def decode_logits(params: Dict[str, np.ndarray], z: np.ndarray, encoded_hw=(3, 8), keep=None):
    """vae/models.py:258-266.  Returns logits [B,H,W,C_t]."""
    d = z @ params["decoder/dense1/kernel"] + params["decoder/dense1/bias"]
    a = d.reshape(z.shape[0], encoded_hw[0], encoded_hw[1], 256)
    if keep is not None:
        keep["dense1"] = a
    for name in ("deconv1", "deconv2", "deconv3"):
        pre = conv_scatter(a, params["decoder/%s/kernel" % name]) + params["decoder/%s/bias" % name]
        a = np.maximum(pre, 0.0)
        if keep is not None:
            keep[name] = a
            keep["pre/" + name] = pre
    logits = conv_scatter(a, params["decoder/deconv4/kernel"]) + params["decoder/deconv4/bias"]
    return logits


def decode(params, z, encoded_hw=(3, 8)):
    """VAE.generate_from_latent (vae/models.py:188-191): sigmoid, flattened [B, H*W*C_t]."""
    logits = decode_logits(params, z, encoded_hw)
    return sigmoid(logits).reshape(z.shape[0], -1)


def _encoded_hw(x):
    h, w = x.shape[1:3]
    for _ in range(4):
        h, w = (h - 4) // 2 + 1, (w - 4) // 2 + 1
    return h, w


def recon_elem(loss_type: str, y, logits):
    """Per-element reconstruction loss and its derivative w.r.t. the logits.
    vae/models.py:11-22; tf.nn.sigmoid_cross_entropy_with_logits = max(x,0) - x*y + log1p(exp(-|x|))."""
    s = sigmoid(logits)
    if loss_type == "mse":
        diff = y - s
        return diff * diff, -2.0 * diff * s * (1.0 - s)
    if loss_type == "bce":
        return np.maximum(logits, 0.0) - logits * y + np.log1p(np.exp(-np.abs(logits))), s - y
    if loss_type == "bce_v2":
        e = 1e-10
        val = -(y * np.log(e + s) + (1.0 - y) * np.log(e + 1.0 - s))
        dval_ds = -(y / (e + s) - (1.0 - y) / (e + 1.0 - s))
        return val, dval_ds * s * (1.0 - s)
    raise ValueError(loss_type)


def verify_range(t):
    """vae/models.py:24-30, :89-90 -- inputs and targets must lie in [0, 1]."""
    if t.size and (t.min() < 0.0 or t.max() > 1.0):
        raise ValueError("verify_range failed: min=%r max=%r" % (t.min(), t.max()))


def loss_and_grads(params, x, y, eps, loss_type="mse", beta=1.0, kl_tolerance=0.0,
                   want_grads=True, dtype=np.float64, relu_masks=None):
    """Forward + loss (+ reverse-mode gradients of ``loss = recon + beta*kl`` w.r.t. every variable).

    x [B,H,W,3], y [B,H,W,C_t] in [0,1]; eps [B,z] standard-normal draws (the TF Philox stream
    cannot be reproduced, so the noise is an input).  Returns a dict with mean, logvar, z, logits,
    recon, kl, loss and (if want_grads) grads{name: array}.

    relu_masks (optional {layer: bool array}): ReLU activity pattern to use in the BACKWARD pass instead of
    this run's own (pre > 0).  A finite-precision implementation and float64 disagree on the sign of the few
    pre-activations that are ~0, and one such flip moves every downstream gradient by ~1/sqrt(#positions); the
    GPU tests pass the device's own pattern here so that gradients are compared like for like, and check
    separately that the two patterns differ only where |pre-activation| is negligible.  The dict returned also
    carries "relu_pre" (float64 pre-activations) for that check.
    """
    p = {k: np.asarray(v, dtype) for k, v in params.items()}
    x = np.asarray(x, dtype); y = np.asarray(y, dtype); eps = np.asarray(eps, dtype)
    verify_range(x); verify_range(y)
    bsz = x.shape[0]
    zdim = p["mean/bias"].shape[0]
    keep = {}
    mean, logvar = encode(p, x, keep)
    std = np.exp(0.5 * logvar)
    z = mean + eps * std
    ehw = _encoded_hw(x)
    logits = decode_logits(p, z, ehw, keep)
    lflat = logits.reshape(bsz, -1)
    yflat = y.reshape(bsz, -1)
    elem, dlogit = recon_elem(loss_type, yflat, lflat)
    recon = elem.sum(axis=1).mean()
    kl_rows = -0.5 * np.sum(1.0 + logvar - mean * mean - np.exp(logvar), axis=1)
    kl_active = np.ones(bsz, dtype=bool)
    if kl_tolerance > 0:
        floor = kl_tolerance * zdim
        kl_active = kl_rows >= floor           # tf.maximum: gradient flows to kl_rows when it is the max
        kl_rows = np.maximum(kl_rows, floor)
    kl = kl_rows.mean()
    out = dict(mean=mean, logvar=logvar, z=z, logits=logits, recon=recon, kl=kl, loss=recon + beta * kl,
               relu_pre={k[4:]: v for k, v in keep.items() if k.startswith("pre/")})
    if not want_grads:
        return out
    act = {}
    for name in ("conv1", "conv2", "conv3", "conv4", "deconv1", "deconv2", "deconv3"):
        act[name] = (keep[name] > 0) if relu_masks is None else np.asarray(relu_masks[name], bool).reshape(keep[name].shape)

    g = {}
    # ---- decoder backward
    glog = (dlogit / bsz).reshape(logits.shape)
    w = p["decoder/deconv4/kernel"]
    g["decoder/deconv4/kernel"] = conv_wgrad(glog, keep["deconv3"], w.shape[0])
    g["decoder/deconv4/bias"] = glog.sum(axis=(0, 1, 2))
    ga = conv_gather(glog, w) * act["deconv3"]
    for name, below in (("deconv3", "deconv2"), ("deconv2", "deconv1"), ("deconv1", "dense1")):
        w = p["decoder/%s/kernel" % name]
        g["decoder/%s/kernel" % name] = conv_wgrad(ga, keep[below], w.shape[0])
        g["decoder/%s/bias" % name] = ga.sum(axis=(0, 1, 2))
        ga = conv_gather(ga, w)
        if below != "dense1":
            ga = ga * act[below]
    gd = ga.reshape(bsz, -1)
    g["decoder/dense1/kernel"] = z.T @ gd
    g["decoder/dense1/bias"] = gd.sum(axis=0)
    gz = gd @ p["decoder/dense1/kernel"].T
    # ---- sampling + KL backward
    klmask = kl_active[:, None].astype(dtype)
    gmean = gz + (beta / bsz) * mean * klmask
    glogvar = gz * (0.5 * eps * std) + (beta / bsz) * 0.5 * (np.exp(logvar) - 1.0) * klmask
    flat = keep["conv4"].reshape(bsz, -1)
    g["mean/kernel"] = flat.T @ gmean
    g["mean/bias"] = gmean.sum(axis=0)
    g["logstd_sqare/kernel"] = flat.T @ glogvar
    g["logstd_sqare/bias"] = glogvar.sum(axis=0)
    gflat = gmean @ p["mean/kernel"].T + glogvar @ p["logstd_sqare/kernel"].T
    ga = gflat.reshape(keep["conv4"].shape) * act["conv4"]
    # ---- encoder backward
    inputs = {"conv4": keep["conv3"], "conv3": keep["conv2"], "conv2": keep["conv1"], "conv1": x}
    for name in ("conv4", "conv3", "conv2", "conv1"):
        w = p["encoder/%s/kernel" % name]
        src = inputs[name]
        g["encoder/%s/kernel" % name] = conv_wgrad(src, ga, w.shape[0])
        g["encoder/%s/bias" % name] = ga.sum(axis=(0, 1, 2))
        if name != "conv1":      # the reference computes conv1's input gradient too, and discards it
            ga = conv_scatter(ga, w, out_hw=src.shape[1:3]) * act[{"conv4": "conv3", "conv3": "conv2", "conv2": "conv1"}[name]]
    out["grads"] = g
    return out


# ----------------------------------------------------------------------------- optimiser
# The reference's graph holds these as float32 constants (shipped .meta: vae/Adam/beta1 = 0.8999999761581421,
# beta2 = 0.9990000128746033, epsilon = 9.99999993922529e-09, learning_rate = 9.999999747378752e-05; pinned by
# tests/test_oracle.py::test_constants_pinned_to_the_shipped_graphs) and TF's ApplyAdam kernel computes 1 - beta in
# float32, so the float64 restatement uses the float32-ROUNDED values, not the Python literals.
ADAM_BETA1 = float(np.float32(0.9))
ADAM_BETA2 = float(np.float32(0.999))
ADAM_EPS = float(np.float32(1e-8))


def adam_init_state(params, dtype=np.float64):
    return dict(m={k: np.zeros_like(v, dtype=dtype) for k, v in params.items()},
                v={k: np.zeros_like(v, dtype=dtype) for k, v in params.items()},
                beta1_power=ADAM_BETA1, beta2_power=ADAM_BETA2)


def adam_apply(params, grads, state, lr, beta1=ADAM_BETA1, beta2=ADAM_BETA2, eps=ADAM_EPS):
    """TF-1.13 ``ApplyAdam`` (epsilon-hat form), in place:
        alpha = lr * sqrt(1 - beta2_power) / (1 - beta1_power)
        m += (g - m) * (1 - beta1);  v += (g*g - v) * (1 - beta2);  p -= alpha * m / (sqrt(v) + eps)
    then beta{1,2}_power *= beta{1,2} (AdamOptimizer._finish).  The power accumulators start at
    beta1 / beta2 (so the first step uses t = 1)."""
    lr = float(np.float32(lr))                 # the learning-rate Const of the graph is float32
    alpha = lr * np.sqrt(1.0 - state["beta2_power"]) / (1.0 - state["beta1_power"])
    for k in params:
        g = grads[k]
        m = state["m"][k]; v = state["v"][k]
        m += (g - m) * (1.0 - beta1)
        v += (g * g - v) * (1.0 - beta2)
        params[k] -= alpha * m / (np.sqrt(v) + eps)
    state["beta1_power"] *= beta1
    state["beta2_power"] *= beta2


def train_step(params, state, x, y, eps, lr=1e-4, loss_type="mse", beta=1.0, kl_tolerance=0.0, dtype=np.float64):
    """One reference minibatch step (vae/models.py:213-216): params/state updated in place
    (they must already be ``dtype`` arrays).  Returns (recon, kl)."""
    out = loss_and_grads(params, x, y, eps, loss_type, beta, kl_tolerance, True, dtype)
    adam_apply(params, out["grads"], state, lr)
    return out["recon"], out["kl"]


# ----------------------------------------------------------------------------- MlpVAE (vae/models.py:271-299)
def mlp_param_shapes(source_shape=(80, 160, 3), target_channels=3, z_dim=Z_DIM_DEFAULT, encoder_sizes=(512, 256), decoder_sizes=(256, 512)):
    """tf.layers.dense variables in creation order: encoder/dense, encoder/dense_1, mean, logstd_sqare, decoder/dense,
    decoder/dense_1, decoder/dense_2 (build_mlp, vae/models.py:283-296)."""
    n_in = int(np.prod(source_shape))
    n_out = source_shape[0] * source_shape[1] * target_channels
    s = OrderedDict()
    s["encoder/dense/kernel"] = (n_in, encoder_sizes[0]);                  s["encoder/dense/bias"] = (encoder_sizes[0],)
    s["encoder/dense_1/kernel"] = (encoder_sizes[0], encoder_sizes[1]);    s["encoder/dense_1/bias"] = (encoder_sizes[1],)
    s["mean/kernel"] = (encoder_sizes[1], z_dim);                          s["mean/bias"] = (z_dim,)
    s["logstd_sqare/kernel"] = (encoder_sizes[1], z_dim);                  s["logstd_sqare/bias"] = (z_dim,)
    s["decoder/dense/kernel"] = (z_dim, decoder_sizes[0]);                 s["decoder/dense/bias"] = (decoder_sizes[0],)
    s["decoder/dense_1/kernel"] = (decoder_sizes[0], decoder_sizes[1]);    s["decoder/dense_1/bias"] = (decoder_sizes[1],)
    s["decoder/dense_2/kernel"] = (decoder_sizes[1], n_out);               s["decoder/dense_2/bias"] = (n_out,)
    return s


def mlp_glorot_init(seed=0, dtype=np.float32, **kw):
    rng = np.random.RandomState(seed)
    out = OrderedDict()
    for name, shape in mlp_param_shapes(**kw).items():
        if name.endswith("bias"):
            out[name] = np.zeros(shape, dtype)
        else:
            limit = np.sqrt(6.0 / (shape[0] + shape[1]))
            out[name] = rng.uniform(-limit, limit, size=shape).astype(dtype)
    return out


def mlp_loss_and_grads(params, x, y, eps, loss_type="mse", beta=1.0, kl_tolerance=0.0, want_grads=True, dtype=np.float64):
    """Forward + loss (+ hand-derived gradients) of the MlpVAE graph; x [B,H,W,3], y [B,H,W,C_t], eps [B,z]."""
    p = {k: np.asarray(v, dtype) for k, v in params.items()}
    x = np.asarray(x, dtype); y = np.asarray(y, dtype); eps = np.asarray(eps, dtype)
    verify_range(x); verify_range(y)
    b = x.shape[0]
    xf = x.reshape(b, -1); yf = y.reshape(b, -1)
    h1 = np.maximum(xf @ p["encoder/dense/kernel"] + p["encoder/dense/bias"], 0.0)
    h2 = np.maximum(h1 @ p["encoder/dense_1/kernel"] + p["encoder/dense_1/bias"], 0.0)
    mean = h2 @ p["mean/kernel"] + p["mean/bias"]
    logvar = h2 @ p["logstd_sqare/kernel"] + p["logstd_sqare/bias"]
    std = np.exp(0.5 * logvar)
    z = mean + eps * std
    g1 = np.maximum(z @ p["decoder/dense/kernel"] + p["decoder/dense/bias"], 0.0)
    g2 = np.maximum(g1 @ p["decoder/dense_1/kernel"] + p["decoder/dense_1/bias"], 0.0)
    logits = g2 @ p["decoder/dense_2/kernel"] + p["decoder/dense_2/bias"]
    elem, dlogit = recon_elem(loss_type, yf, logits)
    recon = elem.sum(axis=1).mean()
    kl_rows = -0.5 * np.sum(1.0 + logvar - mean * mean - np.exp(logvar), axis=1)
    kl_active = np.ones(b, dtype=bool)
    if kl_tolerance > 0:
        floor = kl_tolerance * mean.shape[1]
        kl_active = kl_rows >= floor
        kl_rows = np.maximum(kl_rows, floor)
    kl = kl_rows.mean()
    out = dict(mean=mean, logvar=logvar, z=z, logits=logits, recon=recon, kl=kl, loss=recon + beta * kl)
    if not want_grads:
        return out
    g = {}
    gl = dlogit / b
    g["decoder/dense_2/kernel"] = g2.T @ gl; g["decoder/dense_2/bias"] = gl.sum(axis=0)
    d = (gl @ p["decoder/dense_2/kernel"].T) * (g2 > 0)
    g["decoder/dense_1/kernel"] = g1.T @ d; g["decoder/dense_1/bias"] = d.sum(axis=0)
    d = (d @ p["decoder/dense_1/kernel"].T) * (g1 > 0)
    g["decoder/dense/kernel"] = z.T @ d; g["decoder/dense/bias"] = d.sum(axis=0)
    gz = d @ p["decoder/dense/kernel"].T
    klmask = kl_active[:, None].astype(dtype)
    gmean = gz + (beta / b) * mean * klmask
    glogvar = gz * (0.5 * eps * std) + (beta / b) * 0.5 * (np.exp(logvar) - 1.0) * klmask
    g["mean/kernel"] = h2.T @ gmean; g["mean/bias"] = gmean.sum(axis=0)
    g["logstd_sqare/kernel"] = h2.T @ glogvar; g["logstd_sqare/bias"] = glogvar.sum(axis=0)
    d = (gmean @ p["mean/kernel"].T + glogvar @ p["logstd_sqare/kernel"].T) * (h2 > 0)
    g["encoder/dense_1/kernel"] = h1.T @ d; g["encoder/dense_1/bias"] = d.sum(axis=0)
    d = (d @ p["encoder/dense_1/kernel"].T) * (h1 > 0)
    g["encoder/dense/kernel"] = xf.T @ d; g["encoder/dense/bias"] = d.sum(axis=0)
    out["grads"] = g
    return out


def mlp_train_step(params, state, x, y, eps, lr=1e-4, loss_type="mse", beta=1.0, kl_tolerance=0.0, dtype=np.float64):
    out = mlp_loss_and_grads(params, x, y, eps, loss_type, beta, kl_tolerance, True, dtype)
    adam_apply(params, out["grads"], state, lr)
    return out["recon"], out["kl"]
