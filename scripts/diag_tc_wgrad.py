"""Probe of the MN-major operand descriptor semantics: out = big^T small through tc_wgrad for each variant."""
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
    for variant in range(4):
        out = torch.zeros(i, j, device="cuda"); part = torch.zeros(2 * i * j, device="cuda")
        _lib.check(lib.cpb_debug_tc_wgrad(tb.data_ptr(), ts.data_ptr(), out.data_ptr(), m, i, j, variant, part.data_ptr(), _lib.current_stream_handle()))
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        print("M=%d I=%d J=%d variant %d (placement %d, swap %d): rel err %.3e" % (m, i, j, variant, variant & 1, variant >> 1, np.linalg.norm(got - ref) / np.linalg.norm(ref)))
