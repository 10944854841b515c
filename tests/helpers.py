"""Shared helpers of the parity tests."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_l2(a, b):
    """norm-wise relative error ||a-b|| / ||b|| (the parity metric)."""
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    den = np.linalg.norm(b)
    return np.linalg.norm(a - b) / den if den > 0 else np.linalg.norm(a - b)


def shipped_vae_weights():
    z = np.load(os.path.join(GOLDEN, "vae_rgb_ckpt232.npz"))
    from oracle.vae_oracle import param_shapes
    return {k: z[k] for k in param_shapes().keys()}, z


def shipped_ppo(prefix="policy"):
    z = np.load(os.path.join(GOLDEN, "ppo_ckpt705.npz"))
    from oracle.ppo_oracle import PPO_TENSORS
    return {k: z["%s/%s" % (prefix, k)] for k in PPO_TENSORS}, z


def committed_frames():
    z = np.load(os.path.join(GOLDEN, "frames_u8.npz"))
    return z["rgb"], z["seg"]


def kat():
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        return json.load(f)


class Box:
    """Minimal stand-in for gym.spaces.Box (the PPO class only needs shape/low/high, ppo.py:38)."""

    def __init__(self, low, high):
        self.low = np.asarray(low, np.float32)
        self.high = np.asarray(high, np.float32)
        self.shape = self.low.shape
