"""Generates the committed golden fixtures from the reference's shipped artefacts.

Run ONCE in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

The GPU box has no /root/reference, so everything the ``-m gpu`` tests, ``smoke()`` and
``bench.py`` need from the reference's fixtures is committed here as small ``.npz``/``.json`` files:

  vae_rgb_ckpt232.npz   the 22 weight tensors of vae/models/rgb_bce_cnn_zdim64_beta1_kl_tolerance0.0_data/
                        checkpoints/model.ckpt-232 (+ beta powers, step_idx, and the Adam m/v slots of two
                        small tensors for the isolated Adam test)
  ppo_ckpt705.npz       policy/* and policy_old/* (13 tensors each) + policy Adam slots + counters of
                        models/pretrained_agent/checkpoints/model.ckpt-705
  frames_u8.npz         128 shipped frames vae/data/rgb/{i}.png and channel 0 of vae/data/segmentation/{i}.png
  kat.json              the reference's logged losses (tfevents) and the float64 oracle's outputs on the
                        committed frames / on BASELINE config 1, so the GPU parity tests have fixed targets
"""
import json
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

from carla_ppo_b200.tf_bundle import BundleReader          # noqa: E402
from oracle import vae_oracle as vo                         # noqa: E402
from oracle import ppo_oracle as po                         # noqa: E402

RGB_DIR = "vae/models/rgb_bce_cnn_zdim64_beta1_kl_tolerance0.0_data"
N_FRAMES = 128


def main():
    # ---- VAE checkpoint
    r = BundleReader(os.path.join(REF, RGB_DIR, "checkpoints/model.ckpt-232"))
    names = list(vo.param_shapes().keys())
    vae = {n: r.get("vae/" + n) for n in names}
    extra = {
        "beta1_power": r.get("vae/beta1_power"), "beta2_power": r.get("vae/beta2_power"),
        "step_idx": r.get("vae/step_idx"),
    }
    for n in ("encoder/conv2/kernel", "mean/bias"):
        extra["adam_m/" + n] = r.get("vae/vae/%s/Adam" % n)
        extra["adam_v/" + n] = r.get("vae/vae/%s/Adam_1" % n)
    np.savez(os.path.join(OUT, "vae_rgb_ckpt232.npz"), **vae, **extra)

    # ---- PPO checkpoint
    r2 = BundleReader(os.path.join(REF, "models/pretrained_agent/checkpoints/model.ckpt-705"))
    ppo = {}
    for n in po.PPO_TENSORS:
        ppo["policy/" + n] = r2.get("policy/" + n)
        ppo["policy_old/" + n] = r2.get("policy_old/" + n)
        ppo["adam_m/" + n] = r2.get("policy/%s/Adam" % n)
        ppo["adam_v/" + n] = r2.get("policy/%s/Adam_1" % n)
    for n in ("beta1_power", "beta2_power", "episode_counter", "train_step_counter", "predict_step_counter"):
        ppo[n] = r2.get(n)
    np.savez(os.path.join(OUT, "ppo_ckpt705.npz"), **ppo)

    # ---- frames
    idx = np.sort(np.random.RandomState(2024).choice(10000, N_FRAMES, replace=False))
    rgb = np.stack([np.asarray(Image.open(os.path.join(REF, "vae/data/rgb/%d.png" % i)))[:, :, :3] for i in idx])
    seg = np.stack([np.asarray(Image.open(os.path.join(REF, "vae/data/segmentation/%d.png" % i)))[:, :, 0] for i in idx])
    assert rgb.dtype == np.uint8 and rgb.shape == (N_FRAMES, 80, 160, 3)
    np.savez_compressed(os.path.join(OUT, "frames_u8.npz"), index=idx, rgb=rgb, seg=seg)

    # ---- logged losses + oracle outputs
    from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
    kat = {"source": "reference tfevents + float64 oracle (oracle/vae_oracle.py)", "logged": {}}
    for split in ("train", "val"):
        ea = EventAccumulator(os.path.join(REF, RGB_DIR, "logs", split), size_guidance={"scalars": 0})
        ea.Reload()
        kat["logged"][split] = {
            tag: [[e.step, float(e.value)] for e in ea.Scalars(tag)[-12:]]
            for tag in ("vae/reconstruction_loss", "vae/kl_loss")}

    x = rgb.astype(np.float32) / 255.0               # vae/train_vae.py:15-18
    eps = np.random.RandomState(7).randn(N_FRAMES, 64)
    out = vo.loss_and_grads(vae, x, x, eps, "bce", want_grads=False)
    kat["rgb232_bce_on_committed_frames"] = {"recon": out["recon"], "kl": out["kl"], "eps_seed": 7}
    out = vo.loss_and_grads(vae, x, x, eps, "mse", want_grads=False)
    kat["rgb232_mse_on_committed_frames"] = {"recon": out["recon"], "kl": out["kl"], "eps_seed": 7}

    # BASELINE config 1: 32 random frames, MSE, shipped weights and glorot(seed 0)
    x1 = np.random.RandomState(0).rand(32, 80, 160, 3).astype(np.float32)
    e1 = np.random.RandomState(1).randn(32, 64)
    cfg1 = {}
    for tag, params in (("shipped", vae), ("glorot0", vo.glorot_init(0))):
        o = vo.loss_and_grads(params, x1, x1, e1, "mse", want_grads=True)
        cfg1[tag] = {"recon": o["recon"], "kl": o["kl"],
                     "mean_l2": float(np.linalg.norm(o["mean"])), "logvar_l2": float(np.linalg.norm(o["logvar"])),
                     "logits_l2": float(np.linalg.norm(o["logits"])),
                     "grad_l2": {k: float(np.linalg.norm(v)) for k, v in o["grads"].items()}}
        np.savez(os.path.join(OUT, "config1_%s.npz" % tag), mean=o["mean"], logvar=o["logvar"],
                 sigmoid_sample=vo.sigmoid(o["logits"])[:2])
    kat["config1"] = cfg1
    with open(os.path.join(OUT, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    print("golden fixtures written to", OUT)
    for fn in sorted(os.listdir(OUT)):
        print("  %-28s %9d bytes" % (fn, os.path.getsize(os.path.join(OUT, fn))))


if __name__ == "__main__":
    main()
