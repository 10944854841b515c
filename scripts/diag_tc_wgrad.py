"""Tensor-core weight-gradient kernel vs float64: out = big^T small through cpb_debug_tc_wgrad."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from carla_ppo_b200 import _lib
lib = _lib.load()
rs = np.random.RandomState(0)
for (m, i, j) in [(4096, 128, 128), (5000, 256, 64), (4096, 128, 32)]:
    big = rs.randn(m, i).astype(np.float32); small = rs.randn(m, j).astype(np.float32)
    ref = big.astype(np.float64).T @ small.astype(np.float64)
    tb, ts = torch.tensor(big, device="cuda"), torch.tensor(small, device="cuda")
    out = torch.zeros(i, j, device="cuda"); part = torch.zeros(2 * i * j, device="cuda")
    _lib.check(lib.cpb_debug_tc_wgrad(tb.data_ptr(), ts.data_ptr(), out.data_ptr(), m, i, j, 0, part.data_ptr(), _lib.current_stream_handle()))
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    print("M=%d I=%d J=%d variant 0: rel err %.3e" % (m, i, j, np.linalg.norm(got - ref) / np.linalg.norm(ref)))
