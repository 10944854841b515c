"""Device time of one ConvVAE train step at a given per-GPU batch, and its split over the library's call sites:
B=1024 python scripts/step_profile.py   (what one rank of a 4-GPU run executes)"""
import ctypes as C, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from carla_ppo_b200 import _lib
from carla_ppo_b200.vae.models import ConvVAE

lib = _lib.load()
B = int(os.environ.get("B", "1024"))
vae = ConvVAE((80, 160, 3), z_dim=64, beta=1.0, learning_rate=1e-4, loss_fn="mse", model_dir=tempfile.mkdtemp(), seed=0)
vae.init_session(init_logging=False)
x = torch.rand(B, 80, 160, 3, device="cuda"); eps = torch.randn(B, 64, device="cuda")
for _ in range(3):
    vae.train_step_device(x, x, eps)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 10
e0.record()
for _ in range(steps):
    vae.train_step_device(x, x, eps)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
lib.cpb_profile_reset(); lib.cpb_profile_enable(1)
for _ in range(3):
    vae.train_step_device(x, x, eps)
torch.cuda.synchronize(); lib.cpb_profile_enable(0)
buf = C.create_string_buffer(1 << 16); n = lib.cpb_profile_report(buf, len(buf))
rows = [ln.split() for ln in buf.raw[:n].decode().splitlines()]
tot = sum(float(r[2]) for r in rows) / 3
print("B=%d: %.3f ms/step (%.0f frames/s); sum of labelled groups %.3f ms" % (B, ms, B / ms * 1e3, tot))
if os.environ.get("OUT"):
    import json
    with open(os.environ["OUT"], "w") as f:
        json.dump({"per_gpu_batch": B, "ms_per_step": ms, "frames_per_s": B / ms * 1e3, "sum_of_groups_ms": tot,
                   "groups_ms_per_step": {r[0]: round(float(r[2]) / 3, 4) for r in sorted(rows, key=lambda r: -float(r[2]))},
                   "how": "scripts/step_profile.py: CUDA events around 10 steps; per-call-site CUDA events (cpb_profile_*) over 3 more steps"}, f, indent=1)
for r in sorted(rows, key=lambda r: -float(r[2]))[:40]:
    print("  %-20s %8.3f ms" % (r[0], float(r[2]) / 3))
