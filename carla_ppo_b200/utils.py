"""Drop-in for the hot-path part of the reference's ``utils.py``: ``compute_gae`` (utils.py:45-50),
evaluated by the float64 backward-scan CUDA kernel behind ``cpb_gae``.

``build_mlp`` / ``create_counter_variable`` / ``create_mean_metrics_from_dict`` (utils.py:25-43) are
TensorFlow graph builders with no meaning outside TF; ``VideoRecorder`` (utils.py:9-23) is the same thin OpenCV
writer (used by run_eval when a video file is requested).
"""
from __future__ import annotations

import numpy as np

from . import _lib


def compute_gae(rewards, values, bootstrap_values, terminals, gamma, lam):
    """-> np.ndarray float64 [T]:  delta_t = r_t + (1 - d_t) gamma V_{t+1} - V_t,
    A_t = delta_t + gamma*lam*A_{t+1} (the accumulation is NOT reset at terminals, exactly like the
    reference's lfilter form)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    r = np.ascontiguousarray(np.asarray(rewards, dtype=np.float64).reshape(-1))
    v = np.ascontiguousarray(np.asarray(values, dtype=np.float64).reshape(-1))
    d = np.ascontiguousarray(np.asarray(terminals, dtype=np.float64).reshape(-1))
    t_len = r.shape[0]
    if t_len == 0:
        return np.zeros(0, np.float64)
    if v.shape[0] != t_len or d.shape[0] != t_len:
        raise ValueError("compute_gae: rewards, values and terminals must have the same length")
    packed = torch.from_numpy(np.concatenate([r, v, d])).cuda()
    out = torch.empty(t_len, dtype=torch.float64, device=packed.device)
    base = packed.data_ptr()
    _lib.check(lib.cpb_gae(base, base + 8 * t_len, float(np.asarray(bootstrap_values).reshape(-1)[0]),
                           base + 16 * t_len, t_len, float(gamma), float(lam), out.data_ptr(), None, None,
                           _lib.current_stream_handle()), "cpb_gae")
    return out.cpu().numpy()


class VideoRecorder:
    """utils.py:9-23: AVI (MPEG) writer for the evaluation episodes; frames are RGB arrays."""

    def __init__(self, filename, frame_size, fps=30):
        import cv2
        self._cv2 = cv2
        self.video_writer = cv2.VideoWriter(filename, cv2.VideoWriter_fourcc(*"MPEG"), int(fps), (frame_size[1], frame_size[0]))

    def add_frame(self, frame):
        self.video_writer.write(self._cv2.cvtColor(frame, self._cv2.COLOR_RGB2BGR))

    def release(self):
        self.video_writer.release()

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass
