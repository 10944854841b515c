"""Rebuilds the two tables of profiles/r2_ncu_summary.md from profiles/r2_ncu_full_raw_{tc,other}.csv (text around them is kept)."""
import csv, os, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")

def load(path):
    rows = list(csv.reader(open(path)))
    return rows, {n: i for i, n in enumerate(rows[0])}

rows, col = load(os.path.join(P, "r2_ncu_full_raw_tc.csv"))
K = {"t": "gpu__time_duration.sum", "tp": "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
     "tcw": "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
     "lsw": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum"}
tg = ["conv2.fwd", "conv3.fwd", "conv4.fwd", "deconv1.fwd", "deconv2.fwd", "deconv3.fwd", "deconv3.dgrad", "deconv2.dgrad", "deconv1.dgrad",
      "conv4.dgrad", "conv3.dgrad", "conv2.dgrad"]
wg = ["deconv3.wgrad", "deconv2.wgrad", "deconv1.wgrad", "conv4.wgrad", "conv3.wgrad", "conv2.wgrad"]
ti = wi = 0
t1 = ["| launch | kernel | time under ncu (ms) | tensor pipe active % | smem wavefronts: tensor core % / LSU % of peak | DRAM read + write (GB) |", "|---|---|---|---|---|---|"]
for r in rows[2:]:
    if "tapgemm" in r[col["Kernel Name"]]:
        nm, k = tg[ti], "tc2_tapgemm<pair>"; ti += 1
    else:
        nm, k = wg[wi], "tc_wgrad"; wi += 1
    f = lambda key: float(r[col[K[key]]])
    t1.append("| %s | %s | %.3f | %.1f | %.1f / %.1f | %.2f + %.2f |" % (nm, k, f("t"), f("tp"), f("tcw"), f("lsw"), f("rd"), f("wr")))
rows, col = load(os.path.join(P, "r2_ncu_full_raw_other.csv"))
t2 = ["| kernel | time under ncu (ms) | DRAM read + write (%s) | fp32 FMA pipe, %% of issue peak |" % rows[1][col["dram__bytes_read.sum"]], "|---|---|---|---|"]
for r in rows[2:]:
    t2.append("| %s | %.3f | %s + %s | %.1f |" % (re.sub(r"^void |unnamed>::|\(.*$", "", r[col["Kernel Name"]])[:40], float(r[col["gpu__time_duration.sum"]]),
                                               r[col["dram__bytes_read.sum"]], r[col["dram__bytes_write.sum"]], float(r[col["sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"]])))
md = open(os.path.join(P, "r2_ncu_summary.md")).read()
a = md.index("## Tensor-core launches") + len("## Tensor-core launches\n\n"); b = md.index("\n\n* `sm__pipe_tensor_cycles_active`")
md = md[:a] + "\n".join(t1) + md[b:]
a = md.index("## SIMT edge / reduction kernels") + len("## SIMT edge / reduction kernels\n\n"); b = md.index("\n\n`edge_wgrad<3>`")
md = md[:a] + "\n".join(t2) + md[b:]
open(os.path.join(P, "r2_ncu_summary.md"), "w").write(md)
print("\n".join(t1[:5])); print("\n".join(t2[:4]))
