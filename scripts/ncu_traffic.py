"""Builds profiles/r2_ncu_traffic.json (what bench.py reports as roofline.traffic) from an `ncu --set full` raw CSV:

    ncu --set full --import-source on --clock-control none -k regex:"tc2_|edge_|deconv4" --launch-skip N -c M \\
        -o gpurun_out/r2_full python scripts/tc_prof.py
    ncu -i gpurun_out/r2_full.ncu-rep --page raw --csv > profiles/r2_ncu_full_raw.csv
    python scripts/ncu_traffic.py profiles/r2_ncu_full_raw.csv

The kernels of one train step appear in launch order; the labelled groups are assigned by that order (the same order
cpb_profile_report uses).  The JSON also stores the hash of the CUDA sources so that bench.py can flag a stale table."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# launch order of the tensor-core kernels inside one cpb_vae_loss_grad (vae_api.cu run_forward_loss / run_backward)
TAPGEMM_ORDER = ["conv2.fwd", "conv3.fwd", "conv4.fwd", "deconv1.fwd", "deconv2.fwd", "deconv3.fwd",
                 "deconv3.dgrad", "deconv2.dgrad", "deconv1.dgrad", "conv4.dgrad", "conv3.dgrad", "conv2.dgrad"]
WGRAD_ORDER = ["deconv3.wgrad", "deconv2.wgrad", "deconv1.wgrad", "conv4.wgrad", "conv3.wgrad", "conv2.wgrad"]


def main(path):
    from bench import kernel_source_hash
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    col = {n: i for i, n in enumerate(hdr)}
    name_i = col["Kernel Name"]
    rd_i, wr_i = col["dram__bytes_read.sum"], col["dram__bytes_write.sum"]
    units = rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def val(r, i):
        return float(r[i].replace(",", "")) * scale.get(units[i], 1.0)
    tg, wg, out = [], [], {}
    for r in rows[2:]:
        if len(r) <= max(rd_i, wr_i):
            continue
        b = val(r, rd_i) + val(r, wr_i)
        if "tapgemm_kernel" in r[name_i]:
            tg.append(b)
        elif "wgrad_kernel" in r[name_i] and "edge" not in r[name_i]:
            wg.append(b)
    for k, b in zip(TAPGEMM_ORDER, tg[-len(TAPGEMM_ORDER):]):
        out[k] = b
    for k, b in zip(WGRAD_ORDER, wg[-len(WGRAD_ORDER):]):
        out[k] = b
    blob = {"bytes_per_launch": out, "source_hash": kernel_source_hash(), "from": os.path.relpath(path, ROOT),
            "what": "dram__bytes_read.sum + dram__bytes_write.sum of the LAST train step in the capture, B=4096"}
    with open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json"), "w") as f:
        json.dump(blob, f, indent=1)
    print(json.dumps(blob, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
