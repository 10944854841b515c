"""GPU diagnostics: per-tensor gradient errors vs the fp64 oracle (and vs a fp32 CPU restatement), Adam
trajectory errors, and a rough timing of the B=4096 step.  Not a test; prints a report."""
import os, sys, time, tempfile
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import rel_l2, shipped_vae_weights
from oracle import vae_oracle as vo, torch_ref as tr
from carla_ppo_b200.vae.models import ConvVAE

tmp = tempfile.mkdtemp()
from carla_ppo_b200 import _lib
MODE = int(os.environ.get("CPB_MATH_MODE", "1"))
_lib.check(_lib.load().cpb_set_math_mode(MODE))
print("math mode", _lib.load().cpb_get_math_mode())
def make(w, **kw):
    v = ConvVAE((80,160,3), z_dim=64, loss_fn=kw.pop("loss","mse"), model_dir=tmp, seed=0, **kw); v.init_session(init_logging=False); v.set_weights(w); return v
def dev(a): return torch.as_tensor(np.ascontiguousarray(a), device="cuda")

def grads_report(tag, w, n):
    x = np.random.RandomState(0).rand(n,80,160,3).astype(np.float32); eps = np.random.RandomState(1).randn(n,64).astype(np.float32)
    vae = make(w)
    vae.loss_grad_device(dev(x), dev(x), dev(eps))
    got = vae.get_grads()
    ref = vo.loss_and_grads(w, x, x, eps, "mse")
    r32 = tr.vae_loss_and_grads(w, x, x, eps, "mse", dtype=torch.float32)
    print("== %s  B=%d   tensor: gpu_err  cpu32_err  |g|" % (tag, n))
    for k, g in ref["grads"].items():
        d = np.abs(got[k].astype(np.float64) - g)
        print("  %-26s %.2e  %.2e  %.2e   max|d|=%.2e at %s" % (k, rel_l2(got[k], g), rel_l2(r32["grads"][k], g), np.linalg.norm(g), d.max(), np.unravel_index(d.argmax(), d.shape)))

if "grads" in sys.argv or len(sys.argv) == 1:
    grads_report("glorot3", vo.glorot_init(3), 7)
    grads_report("shipped", shipped_vae_weights()[0], 16)

if "traj" in sys.argv or len(sys.argv) == 1:
    w = shipped_vae_weights()[0]
    vae = make(w)
    p64 = {k: v.astype(np.float64) for k, v in w.items()}; st = vo.adam_init_state(p64)
    t32 = tr.TorchVAETrainer(w, lr=1e-4)
    rs = np.random.RandomState(11)
    for step in range(3):
        x = rs.rand(8,80,160,3).astype(np.float32); eps = rs.randn(8,64).astype(np.float32)
        vae.train_step_device(dev(x), dev(x), dev(eps))
        vo.train_step(p64, st, x, x, eps, lr=1e-4)
        t32.step(torch.from_numpy(x), torch.from_numpy(x), torch.from_numpy(eps))
        got = vae.get_weights()
        print("== step %d: tensor  gpu_param_err cpu32_param_err | gpu_delta_err cpu32_delta_err" % step)
        for k in p64:
            c32 = t32.p[k].detach().numpy()
            print("  %-26s %.2e %.2e | %.2e %.2e" % (k, rel_l2(got[k], p64[k]), rel_l2(c32, p64[k]),
                  rel_l2(got[k]-w[k], p64[k]-w[k]), rel_l2(c32-w[k], p64[k]-w[k])))

if "time" in sys.argv or len(sys.argv) == 1:
    w = vo.glorot_init(0)
    vae = make(w)
    B = 4096
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    x = torch.rand(B,80,160,3, generator=g, device="cuda"); eps = torch.randn(B,64, generator=g, device="cuda")
    for _ in range(3): vae.train_step_device(x, x, eps)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): vae.train_step_device(x, x, eps)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/5
    print("B=4096 train step: %.2f ms  -> %.0f frames/s  (%.1f TFLOP/s algorithmic)" % (ms, B/ms*1e3, 3.181/ms*1e3))
