"""Does PPO.learn() (configs[2]) get faster as a replayed CUDA graph?  Eager vs graph time, and bitwise equality of the result."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from carla_ppo_b200.ppo import PPO

class Box:
    low = np.array([-1.0, 0.0], np.float32); high = np.array([1.0, 1.0], np.float32); shape = (2,)
pol, old, am, av, pw = bench.shipped_agent()
s_, a_, r_, v_, d_, perms_ = bench.ppo_config3_inputs()

def make():
    p = PPO((67,), Box(), learning_rate=1e-4, value_scale=1.0, entropy_scale=0.01, epsilon=0.2, model_dir=tempfile.mkdtemp(), seed=0)
    p.init_session(init_logging=False); p.set_weights(pol, old, am, av, pw)
    return p
eager, graphed = make(), make()
dev = eager._device
s = torch.from_numpy(s_).to(dev); a = torch.from_numpy(a_).to(dev); r = torch.from_numpy(r_).to(dev)
v = torch.from_numpy(v_.astype(np.float64)).to(dev); d = torch.from_numpy(d_.astype(np.float64)).to(dev)
perms = torch.from_numpy(perms_.astype(np.int32)).to(dev)
call = lambda p: p.learn(s, a, v, r, d, 0.3, num_epochs=4, batch_size=256, perms=perms)
N = 10
def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(N): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / N
for _ in range(3): call(eager); call(graphed)
t_eager = timed(lambda: call(eager))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    call(graphed)
torch.cuda.synchronize()
call(eager)                      # the capture pass does not execute: keep the two instances in step (capture = 1 learn when replayed)
t_graph = timed(g.replay)
g.replay(); torch.cuda.synchronize()
for _ in range(N - 1 + 1): pass
# bring eager to the same number of learn() calls: graphed did 3 + N + 1 replays, eager 3 + N + 1
print("eager %.3f ms  graph %.3f ms per learn()" % (t_eager, t_graph))
print("bitwise equal params:", bool(torch.equal(eager.params, graphed.params)))
