"""Pins of the oracle itself (CPU only): hand-derived backward vs torch autograd, the reference's own
artefacts (shipped checkpoints, frames, logged losses, Adam beta-powers), and -- when /root/reference is
present (build container) -- the reference's own compute_gae source and the full 1000-frame known-answer test."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest

from helpers import committed_frames, kat, rel_l2, shipped_ppo, shipped_vae_weights

REF = "/root/reference"


def test_vae_backward_matches_autograd():
    from oracle import vae_oracle as vo, torch_ref as tr
    p = vo.glorot_init(0)
    rs = np.random.RandomState(0)
    x = rs.rand(3, 80, 160, 3).astype(np.float32)
    eps = rs.randn(3, 64)
    for loss, beta, tol in (("mse", 1.0, 0.0), ("bce", 1.0, 0.0), ("bce_v2", 3.0, 0.4)):
        a = vo.loss_and_grads(p, x, x, eps, loss, beta, tol)
        b = tr.vae_loss_and_grads(p, x, x, eps, loss, beta, tol)
        assert abs(a["recon"] - b["recon"]) < 1e-9 * abs(b["recon"])
        assert abs(a["kl"] - b["kl"]) < 1e-9 * max(abs(b["kl"]), 1)
        for k in a["grads"]:
            assert rel_l2(a["grads"][k], b["grads"][k]) < 1e-10, (loss, k)


def test_vae_segmentation_target_backward():
    from oracle import vae_oracle as vo, torch_ref as tr
    p = vo.glorot_init(1, target_channels=1)
    rs = np.random.RandomState(1)
    x = rs.rand(2, 80, 160, 3); y = rs.rand(2, 80, 160, 1); eps = rs.randn(2, 64)
    a = vo.loss_and_grads(p, x, y, eps, "bce")
    b = tr.vae_loss_and_grads(p, x, y, eps, "bce")
    assert a["logits"].shape == (2, 80, 160, 1)
    for k in a["grads"]:
        assert rel_l2(a["grads"][k], b["grads"][k]) < 1e-10, k


def test_known_answer_committed_frames_vs_logged_losses():
    """KAT-1 on the committed subset: shipped ckpt-232 + 128 shipped frames -> BCE recon within 1 % of the
    reference's logged 22 327-22 415 and KL within 5 % of 96.5 (stochastic z => statistical)."""
    from oracle import vae_oracle as vo
    w, _ = shipped_vae_weights()
    rgb, _ = committed_frames()
    x = rgb.astype(np.float32) / 255.0
    eps = np.random.RandomState(7).randn(len(x), 64)
    out = vo.loss_and_grads(w, x, x, eps, "bce", want_grads=False)
    k = kat()
    assert abs(out["recon"] - k["rgb232_bce_on_committed_frames"]["recon"]) < 1e-6
    logged = [v for _, v in k["logged"]["val"]["vae/reconstruction_loss"]] + [v for _, v in k["logged"]["train"]["vae/reconstruction_loss"]]
    assert min(logged) * 0.99 < out["recon"] < max(logged) * 1.01
    logged_kl = np.mean([v for _, v in k["logged"]["train"]["vae/kl_loss"]])
    assert abs(out["kl"] - logged_kl) / logged_kl < 0.05


def test_shipped_adam_state_is_consistent_with_tf_adam_form():
    """beta2_power = 0.999^(steps+1) with 90 steps/epoch (cross-checks the power-after-step convention), and the
    shipped m/v slots are plausible EMA states (v >= 0, |m| <= sqrt(v)/sqrt(1-beta2)-ish)."""
    _, z = shipped_vae_weights()
    b2p = float(z["beta2_power"]); b1p = float(z["beta1_power"])
    steps = np.log(b2p) / np.log(0.999) - 1
    assert abs(steps - round(steps)) < 0.5 and round(steps) % 90 == 0      # whole epochs of 90 minibatches
    assert b1p == 0.0 or b1p < 1e-30
    v = z["adam_v/encoder/conv2/kernel"]
    assert (v >= 0).all()


def test_ppo_backward_matches_autograd_on_shipped_agent():
    from oracle import ppo_oracle as po, torch_ref as tr
    pol, _ = shipped_ppo("policy")
    old, _ = shipped_ppo("policy_old")
    assert max(np.abs(pol[k] - old[k]).max() for k in pol) < 2e-2          # they differ by one update (last 32-sample minibatch steps)
    rs = np.random.RandomState(0)
    s = rs.randn(48, 67); a = np.clip(rs.randn(48, 2), [-1, 0], [1, 1]); ret = rs.randn(48); adv = rs.randn(48)
    low, high = np.array([-1.0, 0.0]), np.array([1.0, 1.0])
    A = po.loss_and_grads(pol, old, s, a, ret, adv, low, high, 0.2, 1.0, 0.01)
    B = tr.ppo_loss_and_grads(pol, old, s, a, ret, adv, low, high, 0.2, 1.0, 0.01)
    assert abs(A["loss"] - B["loss"]) < 1e-10
    assert 0.5 < A["mean_ratio"] < 2.0
    for k in A["grads"]:
        assert rel_l2(A["grads"][k], B["grads"][k]) < 1e-10, k


def test_gae_lfilter_form_equals_recursion_and_is_not_reset_at_terminals():
    from oracle import ppo_oracle as po
    rs = np.random.RandomState(0)
    r = rs.rand(300); v = rs.randn(300); d = rs.rand(300) < 0.1
    a = po.compute_gae(r, v, 0.3, d, 0.99, 0.95)
    b = po.compute_gae_loop(r, v, 0.3, d, 0.99, 0.95)
    assert a.dtype == np.float64 and np.abs(a - b).max() < 1e-12
    # a reset-at-terminal GAE differs: the reference's does NOT reset
    adv = np.zeros(300); acc = 0.0; vv = np.append(v, 0.3)
    for t in range(299, -1, -1):
        nd = 1.0 - d[t]
        acc = r[t] + nd * 0.99 * vv[t + 1] - vv[t] + 0.99 * 0.95 * nd * acc
        adv[t] = acc
    assert np.abs(adv - a).max() > 1e-3


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_compute_gae_against_the_reference_source_itself():
    """Import the reference's utils.py with tensorflow/cv2 stubbed out and run ITS compute_gae."""
    from oracle import ppo_oracle as po
    saved = {k: sys.modules.get(k) for k in ("tensorflow", "cv2")}
    sys.modules["tensorflow"] = types.SimpleNamespace(tanh=None)
    sys.modules["cv2"] = types.ModuleType("cv2")
    try:
        spec = importlib.util.spec_from_file_location("ref_utils", os.path.join(REF, "utils.py"))
        ref_utils = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref_utils)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    rs = np.random.RandomState(3)
    for T in (1, 17, 2048):
        r = list(rs.rand(T)); v = list(rs.randn(T).astype(np.float32)); d = list(rs.rand(T) < 0.05)
        ref = ref_utils.compute_gae(r, v, np.float32(0.25), d, 0.99, 0.95)
        got = po.compute_gae(r, v, np.float32(0.25), d, 0.99, 0.95)
        assert np.array_equal(ref, got)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_known_answer_full_reference_fixtures():
    """KAT-1 (rgb ckpt-232) and KAT-2 (seg ckpt-255) on 400 shipped frames straight from /root/reference,
    through the TF-bundle reader: losses within 1 % / 5 % of the reference's event files."""
    from PIL import Image
    from carla_ppo_b200.tf_bundle import BundleReader, latest_checkpoint
    from oracle import vae_oracle as vo
    idx = np.random.RandomState(0).choice(10000, 400, replace=False)
    rgb = np.stack([np.asarray(Image.open("%s/vae/data/rgb/%d.png" % (REF, i)))[:, :, :3] for i in idx]).astype(np.float32) / 255
    seg = np.stack([np.asarray(Image.open("%s/vae/data/segmentation/%d.png" % (REF, i)))[:, :, :1] for i in idx]).astype(np.float32) / 12
    eps = np.random.RandomState(1).randn(400, 64)
    for tag, y, ct, recon_range, kl_ref in (("rgb", rgb, 3, (22327.0, 22415.0), 96.5), ("seg", seg, 1, (5792.0, 5923.0), 118.0)):
        ck = latest_checkpoint("%s/vae/models/%s_bce_cnn_zdim64_beta1_kl_tolerance0.0_data/checkpoints" % (REF, tag))
        r = BundleReader(ck)
        w = {n: r.get("vae/" + n) for n in vo.param_shapes(target_channels=ct)}
        out = vo.loss_and_grads(w, rgb, y, eps, "bce", want_grads=False)
        assert recon_range[0] * 0.99 < out["recon"] < recon_range[1] * 1.01, (tag, out["recon"])
        assert abs(out["kl"] - kl_ref) / kl_ref < 0.05, (tag, out["kl"])


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_committed_fixtures_equal_the_shipped_checkpoints():
    from carla_ppo_b200.tf_bundle import BundleReader
    w, z = shipped_vae_weights()
    r = BundleReader("%s/vae/models/rgb_bce_cnn_zdim64_beta1_kl_tolerance0.0_data/checkpoints/model.ckpt-232" % REF)
    for k in w:
        assert np.array_equal(w[k], r.get("vae/" + k))
    pol, _ = shipped_ppo("policy")
    r2 = BundleReader("%s/models/pretrained_agent/checkpoints/model.ckpt-705" % REF)
    for k in pol:
        assert np.array_equal(pol[k], r2.get("policy/" + k))
    assert int(r2.get("episode_counter")) == 705
