"""Per-phase cycle accounting of the tensor-core tap-GEMM (instrumented instantiation): CPB_TC_DEBUG=16 python scripts/tc_prof.py"""
import os, sys, tempfile
os.environ.setdefault("CPB_TC_DEBUG", "16")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from carla_ppo_b200.vae.models import ConvVAE

B = int(os.environ.get("B", "4096"))
vae = ConvVAE((80, 160, 3), z_dim=64, beta=1.0, learning_rate=1e-4, loss_fn="mse", model_dir=tempfile.mkdtemp(), seed=0)
vae.init_session(init_logging=False)
x = torch.rand(B, 80, 160, 3, device="cuda")
eps = torch.randn(B, 64, device="cuda")
vae.train_step_device(x, x, eps)
torch.cuda.synchronize()
print("=== step 2", flush=True)
vae.train_step_device(x, x, eps)
torch.cuda.synchronize()
