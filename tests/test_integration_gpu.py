"""GPU tests of the glue either side of the hot path: vae_common (reference vae_common.py:6-62), and -- further down --
the re-authored train.py / run_eval.py loop over the offline replay environment."""
import os
import types

import numpy as np
import pytest

from helpers import committed_frames, rel_l2, shipped_vae_weights

pytestmark = pytest.mark.gpu

RGB_DIR = "rgb_bce_cnn_zdim64_beta1_kl_tolerance0.0_data"


def lay_out_shipped_vae(root):
    """Writes the shipped rgb checkpoint-232 (committed golden npz) as a TF-V2 tensor bundle under the reference's
    directory convention vae/models/<name>/checkpoints/model.ckpt-232.* + the ``checkpoint`` state file."""
    from carla_ppo_b200.tf_bundle import write_bundle
    w, z = shipped_vae_weights()
    ck = os.path.join(root, "vae", "models", RGB_DIR, "checkpoints")
    os.makedirs(ck)
    blob = {"vae/" + k: v for k, v in w.items()}
    blob["vae/step_idx"] = np.int32(232)
    blob["vae/beta1_power"] = np.float32(z["beta1_power"]); blob["vae/beta2_power"] = np.float32(z["beta2_power"])
    write_bundle(os.path.join(ck, "model.ckpt-232"), blob)
    with open(os.path.join(ck, "checkpoint"), "w") as f:
        f.write('model_checkpoint_path: "model.ckpt-232"\nall_model_checkpoint_paths: "model.ckpt-232"\n')
    return os.path.join(root, "vae", "models", RGB_DIR)


class FakeVehicle:
    def __init__(self, steer, throttle, speed):
        self.control = types.SimpleNamespace(steer=steer, throttle=throttle)
        self._speed = speed

    def get_speed(self):
        return self._speed

    def get_forward_vector(self):
        return types.SimpleNamespace(x=0.6, y=0.8, z=0.0)


def test_vae_common_load_and_encode_state(tmp_path):
    """load_vae parses z_dim / model type / target depth from the directory name and restores the TF-V2 bundle;
    create_encode_state_fn(env) = [VAE mean of the frame | steer, throttle, speed (| forward vector)], float64, for both
    uint8 observations and preprocess_frame()'d float observations (reference vae_common.py:6-62)."""
    from carla_ppo_b200 import vae_common
    from oracle import vae_oracle as vo
    model_dir = lay_out_shipped_vae(str(tmp_path))
    vae = vae_common.load_vae(model_dir, z_dim=None, model_type=None)
    assert vae.z_dim == 64 and vae.target_shape == (80, 160, 3) and vae.training is False
    assert vae.get_step_idx() == 232
    w = shipped_vae_weights()[0]
    got = vae.get_weights()
    assert all(np.array_equal(got[k], w[k]) for k in w)

    rgb, _ = committed_frames()
    env = types.SimpleNamespace(observation=rgb[3], vehicle=FakeVehicle(0.25, 0.5, 7.5))
    fn = vae_common.create_encode_state_fn(vae, ["steer", "throttle", "speed"])
    state = fn(env)
    assert state.shape == (67,) and state.dtype == np.float64          # np.append upcasts (vae_common.py:61)
    p64 = {k: v.astype(np.float64) for k, v in w.items()}
    mu, _ = vo.encode(p64, vae_common.preprocess_frame(rgb[3:4]).astype(np.float64))
    assert rel_l2(state[:64], mu[0]) < 1e-5
    assert np.array_equal(state[64:], [0.25, 0.5, 7.5])
    env_f = types.SimpleNamespace(observation=vae_common.preprocess_frame(rgb[3]), vehicle=env.vehicle)
    assert rel_l2(fn(env_f)[:64], mu[0]) < 1e-5                        # float frames take the same path
    fn4 = vae_common.create_encode_state_fn(vae, ["steer", "throttle", "speed", "orientation"])
    s4 = fn4(env)
    assert s4.shape == (70,) and np.allclose(s4[67:], [0.6, 0.8, 0.0])
    with pytest.raises(Exception, match="Failed to load VAE"):
        vae_common.load_vae(str(tmp_path / "vae" / "models" / "seg_bce_cnn_zdim64_beta1_kl_tolerance0.0_data"))


def test_checkpoints_written_in_tf_format_round_trip(tmp_path):
    """save(tf_format=True) writes a TF-V2 bundle (what the reference's saver.restore reads); load_latest_checkpoint
    restores weights, Adam slots, beta powers and step_idx from it."""
    from carla_ppo_b200.vae.models import ConvVAE
    from carla_ppo_b200.tf_bundle import BundleReader, verify_bundle_crcs
    w = shipped_vae_weights()[0]
    vae = ConvVAE((80, 160, 3), z_dim=64, loss_fn="mse", model_dir=str(tmp_path / "m"), seed=0)
    vae.init_session(init_logging=False)
    vae.set_weights(w)
    rgb, _ = committed_frames()
    vae.train_step(rgb[:4], rgb[:4], np.random.RandomState(0).randn(4, 64).astype(np.float32))
    vae.step_idx = 5
    vae.save(tf_format=True)
    prefix = os.path.join(vae.checkpoint_dir, "model.ckpt-5")
    assert verify_bundle_crcs(prefix) == 22 * 3 + 3
    names = BundleReader(prefix).keys()
    assert "vae/vae/encoder/conv1/kernel/Adam_1" in names and "vae/beta2_power" in names
    b = ConvVAE((80, 160, 3), z_dim=64, loss_fn="mse", model_dir=str(tmp_path / "m"), seed=1)
    b.init_session(init_logging=False)
    assert b.load_latest_checkpoint() is True and b.get_step_idx() == 5
    assert bool((b.params == vae.params).all()) and bool((b.adam_m == vae.adam_m).all()) and bool((b.adam_v == vae.adam_v).all())
    assert bool((b.adam_powers == vae.adam_powers).all())
