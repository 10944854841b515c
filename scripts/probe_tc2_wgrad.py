"""Descriptor probe of the MN-major tc2 weight-gradient kernel: out = big^T small through cpb_debug_tc_wgrad with
variant 0 (LBO = column-group stride, SBO = 8-row stride), 1 (swapped) and 32 (round-1 register-path kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from carla_ppo_b200 import _lib
lib = _lib.load()

def run(m, i, j, variant):
    rs = np.random.RandomState(m + i + j)
    big = rs.randn(m, i).astype(np.float32); small = rs.randn(m, j).astype(np.float32)
    tb, ts = torch.tensor(big, device="cuda"), torch.tensor(small, device="cuda")
    out = torch.full((i, j), float("nan"), device="cuda")
    partial = torch.empty(2 * i * j, device="cuda")
    st = lib.cpb_debug_tc_wgrad(tb.data_ptr(), ts.data_ptr(), out.data_ptr(), m, i, j, variant, partial.data_ptr(), _lib.current_stream_handle())
    if st < 0:
        return "status %d: %s" % (st, lib.cpb_last_error().decode())
    try:
        torch.cuda.synchronize()
    except Exception as e:
        return "CUDA error: %s" % e
    ref = big.astype(np.float64).T @ small.astype(np.float64)
    got = out.cpu().numpy()
    return "rel err %.3e  finite %s" % (np.linalg.norm(got - ref) / np.linalg.norm(ref), bool(np.isfinite(got).all()))

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for (m, i, j) in [(4096, 128, 128), (5000, 256, 64), (4100, 128, 32), (1031, 384, 256)]:
    print("variant %d  m=%d i=%d j=%d: %s" % (variant, m, i, j, run(m, i, j, variant)), flush=True)
