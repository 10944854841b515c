"""Runs PPO.learn() (BASELINE config 3 shape) a few times; used under ncu for the per-kernel launch list."""
import os, sys, tempfile, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from carla_ppo_b200.ppo import PPO
class Box:
    low = np.array([-1.0, 0.0], np.float32); high = np.array([1.0, 1.0], np.float32); shape = (2,)
ppo = PPO((67,), Box(), learning_rate=1e-4, value_scale=1.0, model_dir=tempfile.mkdtemp(), seed=0); ppo.init_session(init_logging=False)
T, E, B = 2048, 4, 256
rs = np.random.RandomState(0); dev = ppo._device
s = torch.from_numpy(rs.randn(T, 67).astype(np.float32)).to(dev); a = torch.from_numpy(np.clip(rs.randn(T, 2), Box.low, Box.high).astype(np.float32)).to(dev)
r = torch.from_numpy(rs.rand(T)).to(dev); v = torch.from_numpy(rs.randn(T)).to(dev); d = torch.zeros(T, dtype=torch.float64, device=dev); d[-1] = 1
perms = torch.from_numpy(np.stack([np.random.RandomState(e).permutation(T) for e in range(E)]).astype(np.int32)).to(dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for i in range(n):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ppo.learn(s, a, v, r, d, 0.3, num_epochs=E, batch_size=B, perms=perms)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("learn %d: host enqueue %.2f ms, total %.2f ms" % (i, (t1 - t0) * 1e3, (t2 - t0) * 1e3))
