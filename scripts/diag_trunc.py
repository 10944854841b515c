"""Does tcgen05 kind::tf32 truncate its fp32 inputs?  Run once with CPB_TC_DEBUG=0 and once with 32 (A "hi" tile
stored as the raw fp32 value); identical outputs mean the 13 low mantissa bits are ignored by the tensor core."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from carla_ppo_b200 import _lib
lib = _lib.load()
rs = np.random.RandomState(0)
m, n, k = 512, 128, 1024
a = rs.randn(m, k).astype(np.float32); bt = rs.randn(n, k).astype(np.float32)
ta = torch.tensor(a, device="cuda"); tb = torch.tensor(bt, device="cuda")
d = torch.empty(m, n, device="cuda"); sc = torch.empty(2 * n * k + m * k, device="cuda")
_lib.check(lib.cpb_debug_tc_gemm(ta.data_ptr(), tb.data_ptr(), d.data_ptr(), m, n, k, sc.data_ptr(), _lib.current_stream_handle()))
torch.cuda.synchronize()
out = d.cpu().numpy()
ref = a.astype(np.float64) @ bt.astype(np.float64).T
print("flags", os.environ.get("CPB_TC_DEBUG", "0"), "rel err %.3e" % (np.linalg.norm(out - ref) / np.linalg.norm(ref)))
np.save(sys.argv[1], out)
if len(sys.argv) > 2:
    other = np.load(sys.argv[2])
    print("bitwise identical to", sys.argv[2], ":", bool(np.array_equal(out.view(np.uint32), other.view(np.uint32))),
          " max abs diff %.3e" % np.abs(out - other).max())
