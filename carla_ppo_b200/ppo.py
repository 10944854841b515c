"""Drop-in for the reference's ``ppo.py``: the ``PPO`` class with the same constructor, methods and
attributes, backed by libcarla_ppo_b200.so instead of a TensorFlow session.

Reference surface mirrored (paths relative to the reference repo root):
  * ``PPO.__init__`` hyper-parameters, dirs                   ppo.py:73-190
  * ``init_session / save / load_latest_checkpoint``          ppo.py:192-216
  * ``train`` (ONE minibatch Adam step)                       ppo.py:218-229
  * ``predict`` (greedy / sampled+clipped action, value)      ppo.py:231-251
  * counters, summaries, ``update_old_policy``                ppo.py:253-276

Additive entry point: ``learn(...)`` = the driver's whole update block (train.py:171-207: GAE, returns,
advantage normalisation, theta_old <- theta, epochs x shuffled minibatches) in one C call with no host
round trips.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import Dict, Optional

import numpy as np

from . import _lib
from ._lib import CpbError, PpoConfig

ADAM_BETA1, ADAM_BETA2, ADAM_EPS = 0.9, 0.999, 1e-8
_METRIC_NAMES = ("train_loss/policy", "train_loss/value", "train_loss/entropy", "train_loss/loss", "train/prob_ratio")


class PPO:
    def __init__(self, input_shape, action_space, learning_rate=3e-4, lr_decay=0.998, epsilon=0.2,
                 value_scale=0.5, entropy_scale=0.01, initial_std=0.4, model_dir="./", seed=None, device=None):
        input_shape = tuple(int(v) for v in np.atleast_1d(input_shape))
        if len(input_shape) != 1:
            raise ValueError("PPO expects a flat state vector (reference train.py:85 builds [z_dim + measurements])")
        self.input_shape = input_shape
        self.state_dim = input_shape[0]
        self.num_actions = int(action_space.shape[0])
        if self.num_actions > 4:
            raise ValueError("at most 4 action dimensions are supported")
        self.action_low = np.broadcast_to(np.asarray(action_space.low, np.float32), (self.num_actions,)).copy()
        self.action_high = np.broadcast_to(np.asarray(action_space.high, np.float32), (self.num_actions,)).copy()
        self.base_learning_rate = float(learning_rate)
        self.lr_decay = float(lr_decay)
        self.epsilon = float(epsilon)
        self.value_scale = float(value_scale)
        self.entropy_scale = float(entropy_scale)
        self.initial_std = float(initial_std)
        self._seed = seed
        self._device = device

        self.model_dir = model_dir
        self.checkpoint_dir = "{}/checkpoints/".format(self.model_dir)
        self.log_dir = "{}/logs/".format(self.model_dir)
        self.video_dir = "{}/videos/".format(self.model_dir)
        self.dirs = [self.checkpoint_dir, self.log_dir, self.video_dir]
        for d in self.dirs:
            os.makedirs(d, exist_ok=True)

        self.train_step_counter = 0
        self.predict_step_counter = 0
        self.episode_counter = 0
        self.sess = None
        self.train_writer = None
        self._ws = None
        self._pending_metrics = []

    # ------------------------------------------------------------------ session / state
    def _cfg(self):
        cfg = PpoConfig()
        cfg.state_dim, cfg.num_actions, cfg.hidden1, cfg.hidden2 = self.state_dim, self.num_actions, 500, 300
        for k in range(4):
            cfg.action_low[k] = float(self.action_low[k]) if k < self.num_actions else 0.0
            cfg.action_high[k] = float(self.action_high[k]) if k < self.num_actions else 0.0
        cfg.epsilon, cfg.value_scale, cfg.entropy_scale = self.epsilon, self.value_scale, self.entropy_scale
        return cfg

    def init_session(self, sess=None, init_logging=True):
        torch = _lib.require_cuda()
        lib = _lib.load()
        self._torch, self._libh = torch, lib
        if self._device is None:
            self._device = torch.device("cuda", torch.cuda.current_device())
        self._device = torch.device(self._device)
        if self._device.index is None:
            self._device = torch.device("cuda", torch.cuda.current_device())
        dev = self._device
        self._c = self._cfg()
        n = lib.cpb_ppo_num_tensors()
        offs = (C.c_int64 * n)(); sizes = (C.c_int64 * n)(); shapes = (C.c_int32 * (2 * n))()
        total = C.c_int64()
        _lib.check(lib.cpb_ppo_layout(C.byref(self._c), offs, sizes, shapes, C.byref(total)), "cpb_ppo_layout")
        self._names = [lib.cpb_ppo_tensor_name(i).decode() for i in range(n)]
        self._offsets = {self._names[i]: int(offs[i]) for i in range(n)}
        self._shapes = {self._names[i]: tuple(int(s) for s in shapes[2 * i:2 * i + 2] if s > 0) for i in range(n)}
        self._total = int(total.value)
        z = lambda: torch.zeros(self._total, dtype=torch.float32, device=dev)
        self.params, self.params_old, self.grads, self.adam_m, self.adam_v = z(), z(), z(), z(), z()
        self.adam_powers = torch.tensor([ADAM_BETA1, ADAM_BETA2], dtype=torch.float32, device=dev)
        self._lr_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        self._rng = np.random.RandomState(self._seed)
        w = self._initial_weights()
        self.set_weights(w, w)
        self._sync_lr()
        self.sess = self
        if init_logging:
            try:
                from torch.utils.tensorboard import SummaryWriter
                self.train_writer = SummaryWriter(self.log_dir)
            except Exception as e:
                print("carla_ppo_b200: TensorBoard logging disabled (%s)" % e)

    def _initial_weights(self) -> Dict[str, np.ndarray]:
        """tf.layers.dense defaults (glorot uniform / zeros); action_mean kernel = variance_scaling(0.1)
        truncated normal (ppo.py:44-47); action_logstd = log(initial_std) (ppo.py:49)."""
        rng = np.random.RandomState(self._seed if self._seed is not None else np.random.randint(0, 2 ** 31 - 1))
        out = {}
        for name in self._names:
            shape = self._shapes[name]
            if name == "action_logstd":
                out[name] = np.full(shape, np.log(self.initial_std), np.float32)
            elif name.endswith("bias"):
                out[name] = np.zeros(shape, np.float32)
            elif name == "action_mean/kernel":
                std = np.sqrt(0.1 / shape[0]) / 0.87962566103423978
                t = rng.randn(*shape)
                bad = np.abs(t) > 2
                while bad.any():
                    t[bad] = rng.randn(int(bad.sum()))
                    bad = np.abs(t) > 2
                out[name] = (t * std).astype(np.float32)
            else:
                limit = np.sqrt(6.0 / (shape[0] + shape[1]))
                out[name] = rng.uniform(-limit, limit, size=shape).astype(np.float32)
        return out

    def _flatten(self, weights) -> np.ndarray:
        host = np.zeros(self._total, np.float32)
        for name in self._names:
            w = np.asarray(weights[name], np.float32).reshape(self._shapes[name])
            o = self._offsets[name]
            host[o:o + w.size] = w.ravel()
        return host

    def _unflatten(self, flat) -> Dict[str, np.ndarray]:
        host = flat.detach().cpu().numpy()
        return {n: host[self._offsets[n]:self._offsets[n] + int(np.prod(self._shapes[n]))].reshape(self._shapes[n]).copy()
                for n in self._names}

    def set_weights(self, policy, policy_old=None, adam_m=None, adam_v=None, powers=None):
        torch = self._torch
        self.params.copy_(torch.from_numpy(self._flatten(policy)))
        if policy_old is not None:
            self.params_old.copy_(torch.from_numpy(self._flatten(policy_old)))
        if adam_m is not None:
            self.adam_m.copy_(torch.from_numpy(self._flatten(adam_m)))
        if adam_v is not None:
            self.adam_v.copy_(torch.from_numpy(self._flatten(adam_v)))
        if powers is not None:
            self.adam_powers.copy_(torch.tensor([float(powers[0]), float(powers[1])], dtype=torch.float32))

    def get_weights(self):
        return self._unflatten(self.params)

    def get_old_weights(self):
        return self._unflatten(self.params_old)

    def get_grads(self):
        return self._unflatten(self.grads)

    @property
    def learning_rate(self):
        """exponential_decay(learning_rate, episode_counter, 1, lr_decay, staircase) (ppo.py:142)."""
        return np.float32(self.base_learning_rate) * np.float32(self.lr_decay) ** np.float32(self.episode_counter)

    def _sync_lr(self):
        self._lr_dev.fill_(float(self.learning_rate))

    def _require_session(self):
        if self.sess is None:
            raise CpbError("init_session() has not been called")

    def _stream(self):
        return _lib.current_stream_handle(self._device)

    def _call(self, name, *args):
        """C entry point with this model's device current (the library launches on the CURRENT CUDA device)."""
        with self._torch.cuda.device(self._device):
            return _lib.check(getattr(self._libh, name)(*args), name)

    def _workspace(self, max_batch, horizon=0):
        need = self._libh.cpb_ppo_workspace_bytes(C.byref(self._c), int(max_batch), int(horizon))
        _lib.check(need, "cpb_ppo_workspace_bytes")
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = self._torch.empty(int(need), dtype=self._torch.uint8, device=self._device)
        return self._ws

    def _dev(self, a, dtype):
        torch = self._torch
        if isinstance(a, torch.Tensor):
            return a.to(self._device, dtype).contiguous()
        np_dtype = {torch.float32: np.float32, torch.float64: np.float64, torch.int32: np.int32}[dtype]
        return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np_dtype))).to(self._device)

    # ------------------------------------------------------------------ checkpoints
    def save(self, tf_format=False):
        """.npz checkpoint + ``checkpoint`` state file (ppo.py:202-205); ``tf_format=True`` writes a TF-V2 tensor bundle
        with the reference's variable names instead (readable by the reference's ``saver.restore``)."""
        self._require_session()
        step = int(self.episode_counter)
        prefix = os.path.join(self.checkpoint_dir, "model.ckpt-%d" % step)
        blob = {}
        for k, v in self.get_weights().items():
            blob["policy/" + k] = v
        for k, v in self.get_old_weights().items():
            blob["policy_old/" + k] = v
        for k, v in self._unflatten(self.adam_m).items():
            blob["policy/%s/Adam" % k] = v
        for k, v in self._unflatten(self.adam_v).items():
            blob["policy/%s/Adam_1" % k] = v
        pw = self.adam_powers.cpu().numpy()
        blob["beta1_power"], blob["beta2_power"] = pw[0], pw[1]
        blob["episode_counter"] = np.int32(self.episode_counter)
        blob["train_step_counter"] = np.int32(self.train_step_counter)
        blob["predict_step_counter"] = np.int32(self.predict_step_counter)
        if tf_format:
            from .tf_bundle import write_bundle
            write_bundle(prefix, {k: np.asarray(v) for k, v in blob.items()})
        else:
            np.savez(prefix + ".npz", **blob)
        state = os.path.join(self.checkpoint_dir, "checkpoint")
        kept = []
        if os.path.isfile(state):
            with open(state) as f:
                kept = re.findall(r'^all_model_checkpoint_paths:\s*"(.*)"', f.read(), re.M)
        name = os.path.basename(prefix)
        kept = [k for k in kept if k != name] + [name]
        for old in kept[:-5]:
            for ext in (".npz", ".index", ".data-00000-of-00001"):
                try:
                    os.remove(os.path.join(self.checkpoint_dir, old + ext))
                except OSError:
                    pass
        with open(state, "w") as f:
            f.write('model_checkpoint_path: "%s"\n' % name)
            for k in kept[-5:]:
                f.write('all_model_checkpoint_paths: "%s"\n' % k)
        print("Model checkpoint saved to {}".format(prefix))

    def load_latest_checkpoint(self):
        """True / False (restore raised) / None (no checkpoint), like ppo.py:207-216."""
        self._require_session()
        from .tf_bundle import BundleReader
        state = os.path.join(self.checkpoint_dir, "checkpoint")
        if not os.path.isfile(state):
            return None
        with open(state) as f:
            m = re.search(r'^model_checkpoint_path:\s*"(.*)"', f.read(), re.M)
        if not m:
            return None
        prefix = m.group(1)
        if not os.path.isabs(prefix):
            prefix = os.path.join(self.checkpoint_dir, prefix)
        try:
            if os.path.isfile(prefix + ".npz"):
                blob = dict(np.load(prefix + ".npz"))
            elif os.path.isfile(prefix + ".index"):
                blob = BundleReader(prefix).all()
            else:
                return None
            self.load_blob(blob)
            print("Model checkpoint restored from {}".format(prefix))
            return True
        except Exception as e:
            print(e)
            return False

    def load_blob(self, blob):
        pol = {n: blob["policy/" + n] for n in self._names}
        old = {n: blob["policy_old/" + n] for n in self._names}
        m_ = v_ = pw = None
        if ("policy/%s/Adam" % self._names[0]) in blob:
            m_ = {n: blob["policy/%s/Adam" % n] for n in self._names}
            v_ = {n: blob["policy/%s/Adam_1" % n] for n in self._names}
            pw = (float(blob["beta1_power"]), float(blob["beta2_power"]))
        self.set_weights(pol, old, m_, v_, pw)
        for attr in ("episode_counter", "train_step_counter", "predict_step_counter"):
            if attr in blob:
                setattr(self, attr, int(blob[attr]))
        self._sync_lr()

    # ------------------------------------------------------------------ hot path
    def train(self, input_states, taken_actions, returns, advantage):
        """ONE minibatch Adam step (ppo.py:218-229).  Metrics stay on the device until the next
        write_episodic_summaries()."""
        self._require_session()
        torch = self._torch
        s = self._dev(input_states, torch.float32).reshape(-1, self.state_dim)
        a = self._dev(taken_actions, torch.float32).reshape(-1, self.num_actions)
        r = self._dev(returns, torch.float32).reshape(-1)
        adv = self._dev(advantage, torch.float32).reshape(-1)
        b = s.shape[0]
        if not (a.shape[0] == r.shape[0] == adv.shape[0] == b):
            raise ValueError("train(): inconsistent batch sizes")
        metrics = torch.empty(5, dtype=torch.float32, device=self._device)
        ws = self._workspace(b)
        self._call("cpb_ppo_train_step", 
            C.byref(self._c), _lib.ptr(self.params), _lib.ptr(self.params_old), _lib.ptr(self.grads),
            _lib.ptr(self.adam_m), _lib.ptr(self.adam_v), _lib.ptr(self.adam_powers), _lib.ptr(self._lr_dev),
            _lib.ptr(s), _lib.ptr(a), _lib.ptr(r), _lib.ptr(adv), None, b, _lib.ptr(metrics), _lib.ptr(ws),
            ws.numel(), self._stream())
        self._pending_metrics.append(metrics)
        self.train_step_counter += 1
        return metrics

    def loss_and_grads(self, input_states, taken_actions, returns, advantage):
        """Loss + gradients only (no Adam): returns (metrics[5] ndarray, {name: grad})."""
        self._require_session()
        torch = self._torch
        s = self._dev(input_states, torch.float32).reshape(-1, self.state_dim)
        a = self._dev(taken_actions, torch.float32).reshape(-1, self.num_actions)
        r = self._dev(returns, torch.float32).reshape(-1)
        adv = self._dev(advantage, torch.float32).reshape(-1)
        b = s.shape[0]
        metrics = torch.empty(5, dtype=torch.float32, device=self._device)
        ws = self._workspace(b)
        self._call("cpb_ppo_loss_grad", 
            C.byref(self._c), _lib.ptr(self.params), _lib.ptr(self.params_old), _lib.ptr(s), _lib.ptr(a), _lib.ptr(r),
            _lib.ptr(adv), None, b, _lib.ptr(self.grads), _lib.ptr(metrics), _lib.ptr(ws), ws.numel(),
            self._stream())
        return metrics.cpu().numpy(), self.get_grads()

    def predict(self, input_states, greedy=False, write_to_summary=False, noise=None):
        """-> (action, value); squeezed when a single state is given (ppo.py:231-251).  ``noise`` (optional
        [B,A] standard-normal draws) makes the sampled action reproducible."""
        self._require_session()
        torch = self._torch
        x = np.asarray(input_states, dtype=np.float32)
        if x.ndim != 2:
            x = x[None]
        b, a_dim = x.shape[0], self.num_actions
        if greedy:
            packed = x
        else:
            eps = self._rng.randn(b, a_dim).astype(np.float32) if noise is None else np.asarray(noise, np.float32).reshape(b, a_dim)
            packed = np.concatenate([x.reshape(-1), eps.reshape(-1)])
        dev = torch.from_numpy(np.ascontiguousarray(packed).reshape(-1)).to(self._device)
        s = dev[:b * self.state_dim]
        nz = None if greedy else dev[b * self.state_dim:]
        out = torch.empty(b * (a_dim + 1), dtype=torch.float32, device=self._device)
        ws = self._workspace(b)
        self._call("cpb_ppo_forward", C.byref(self._c), _lib.ptr(self.params), _lib.ptr(s), b, _lib.ptr(nz),
                                              _lib.ptr(out), _lib.ptr(out[b * a_dim:]), _lib.ptr(ws), ws.numel(),
                                              self._stream())
        host = out.cpu().numpy()
        action, value = host[:b * a_dim].reshape(b, a_dim), host[b * a_dim:]
        if write_to_summary:
            if self.train_writer is not None:
                for i in range(a_dim):
                    self.train_writer.add_scalar("predict_actor/action_%d/sampled_action" % i, float(action[0, i]),
                                                 self.predict_step_counter)
            self.predict_step_counter += 1
        if b == 1:
            return action[0], value[0]
        return action, value

    def update_old_policy(self):
        """theta_old <- theta (ppo.py:147, 275-276)."""
        self._require_session()
        self.params_old.copy_(self.params)

    def learn(self, states, actions, values, rewards, dones, last_value, gamma=0.99, lam=0.95, num_epochs=3,
              batch_size=32, perms=None, return_metrics=False):
        """train.py:171-207 in one C call: compute_gae -> returns -> normalised advantages ->
        update_old_policy -> num_epochs x ceil(T/batch_size) minibatch steps.  ``perms`` ([num_epochs, T]
        index orders) defaults to np.random permutations like the reference's np.random.shuffle."""
        self._require_session()
        torch = self._torch
        s = self._dev(states, torch.float32).reshape(-1, self.state_dim)
        t_len = s.shape[0]
        a = self._dev(actions, torch.float32).reshape(t_len, self.num_actions)
        r = self._dev(rewards, torch.float64).reshape(t_len)
        v = self._dev(values, torch.float64).reshape(t_len)
        d = self._dev(np.asarray(dones, dtype=np.float64) if not isinstance(dones, torch.Tensor) else dones, torch.float64).reshape(t_len)
        if perms is None:
            perms = np.stack([np.random.permutation(t_len) for _ in range(num_epochs)]) if num_epochs else np.zeros((0, t_len))
        if num_epochs == 0:
            p = None                       # nothing to index: the C entry accepts perms == NULL for zero epochs
        elif isinstance(perms, torch.Tensor):
            p = perms.to(self._device, torch.int32).reshape(num_epochs, t_len).contiguous()
        else:
            p = self._dev(np.asarray(perms).reshape(num_epochs, t_len), torch.int32)
        nmb = -(-t_len // batch_size)
        metrics = torch.empty(max(num_epochs * nmb, 1), 5, dtype=torch.float32, device=self._device)
        ws = self._workspace(min(batch_size, t_len), t_len)
        self._call("cpb_ppo_learn", 
            C.byref(self._c), _lib.ptr(self.params), _lib.ptr(self.params_old), _lib.ptr(self.grads),
            _lib.ptr(self.adam_m), _lib.ptr(self.adam_v), _lib.ptr(self.adam_powers), _lib.ptr(self._lr_dev),
            _lib.ptr(s), _lib.ptr(a), _lib.ptr(r), _lib.ptr(v), float(last_value), _lib.ptr(d), t_len, float(gamma),
            float(lam), int(num_epochs), int(batch_size), _lib.ptr(p), _lib.ptr(metrics), _lib.ptr(ws), ws.numel(),
            self._stream())
        self.train_step_counter += num_epochs * nmb
        self._pending_metrics.append(metrics[:num_epochs * nmb])
        if return_metrics:
            return metrics[:num_epochs * nmb].cpu().numpy().reshape(num_epochs * nmb, 5)
        return None

    # ------------------------------------------------------------------ counters / summaries
    def get_episode_idx(self):
        return int(self.episode_counter)

    def get_train_step_idx(self):
        return int(self.train_step_counter)

    def get_predict_step_idx(self):
        return int(self.predict_step_counter)

    def write_value_to_summary(self, summary_name, value, step):
        if self.train_writer is not None:
            self.train_writer.add_scalar(summary_name, float(value), int(step))

    def write_dict_to_summary(self, summary_name, params, step):
        if self.train_writer is not None:
            self.train_writer.add_text(summary_name, "\n".join("%s: %s" % (k, v) for k, v in params.items()), int(step))

    def write_episodic_summaries(self):
        """Episodic means of the per-minibatch metrics, then episode_counter += 1 (ppo.py:271-273) -- which is
        what decays the learning rate."""
        if self._pending_metrics:
            torch = self._torch
            allm = torch.cat([m.reshape(-1, 5) for m in self._pending_metrics]).double().mean(dim=0).cpu().numpy()
            if self.train_writer is not None:
                for name, val in zip(_METRIC_NAMES, allm):
                    self.train_writer.add_scalar(name, float(val), self.episode_counter)
                self.train_writer.add_scalar("train/learning_rate", float(self.learning_rate), self.episode_counter)
            self._pending_metrics = []
        self.episode_counter += 1
        self._sync_lr()
