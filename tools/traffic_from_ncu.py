"""Per-kernel-family DRAM traffic and time from the `ncu --metrics ... --csv` capture of one forward
(tools/ncu_full_forward.sh).  usage: traffic_from_ncu.py forward_metrics.csv out.json
Families are keyed by the C-ABI entry point bench.py reports (its timeline names)."""
import collections
import csv
import json
import sys

FAMILY = [("tc_mlp_kernel<0>", "pwclo_set_conv_tc"), ("tc_mlp_kernel<1>", "pwclo_pointwise_mlp_tc"),
          ("tc_mlp_kernel<2>", "pwclo_cost_volume_1_tc"), ("tc_mlp_kernel<3>", "pwclo_cost_volume_2_tc"),
          ("set_conv_small_kernel", "pwclo_set_conv"), ("fps_", "pwclo_furthest_point_sampling"),
          ("knn_", "pwclo_knn"), ("pose_head_kernel", "pwclo_pose_head"), ("gather_rows3", "pwclo_gather_rows3"),
          ("transpose_cn", "pwclo_transpose")]

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
col = {n: h.index(n) for n in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
per = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= col["Metric Value"]:
        continue
    k = per.setdefault(r[col["ID"]], {"name": r[col["Kernel Name"]]})
    v = float(r[col["Metric Value"]].replace(",", ""))
    u = r[col["Metric Unit"]]
    v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-3, "ns": 1e-6, "ms": 1.0, "msecond": 1.0, "usecond": 1e-3,
          "nsecond": 1e-6}.get(u, 1.0)
    k[r[col["Metric Name"]]] = v
fam = collections.OrderedDict()
for k in per.values():
    name = next((f for pat, f in FAMILY if pat in k["name"]), k["name"])
    f = fam.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "ms": 0.0, "kernels": set(), "tensor": 0.0, "issue": 0.0,
                              "sm": 0.0})
    ms = k.get("gpu__time_duration.sum", 0.0)
    f["launches"] += 1
    f["dram_bytes"] += k.get("dram__bytes_read.sum", 0.0) + k.get("dram__bytes_write.sum", 0.0)
    f["ms"] += ms
    f["tensor"] += ms * k.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)   # duration-weighted
    f["issue"] += ms * k.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0)
    f["sm"] += ms * k.get("sm__throughput.avg.pct_of_peak_sustained_elapsed", 0.0)
    f["kernels"].add(k["name"].split("(")[0])
out = {n: {"launches_in_capture": f["launches"], "dram_bytes_in_capture": f["dram_bytes"],
           "dram_bytes_per_launch": f["dram_bytes"] / f["launches"], "ms_in_capture_under_ncu": round(f["ms"], 4),
           "tensor_pipe_active_pct": round(f["tensor"] / f["ms"], 2) if f["ms"] else 0.0,
           "issue_active_pct": round(f["issue"] / f["ms"], 2) if f["ms"] else 0.0,
           "sm_throughput_pct": round(f["sm"] / f["ms"], 2) if f["ms"] else 0.0,
           "kernels": sorted(f["kernels"])} for n, f in fam.items()}
json.dump({"source": sys.argv[1], "note": "dram__bytes_read.sum + dram__bytes_write.sum of every launch of ONE launch-by-launch forward of 64 frame pairs x "
           "8192 points (tools/ncu_forward.py); tensor-pipe / issue / SM-throughput percentages are duration-weighted means "
           "over the family's launches; knn = presort + all search kernels", "families": out}, open(sys.argv[2], "w"), indent=1)
for n, f in out.items():
    print(f"{n:34s} {f['launches_in_capture']:3d} launches  {f['dram_bytes_in_capture'] / 1e6:9.2f} MB  {f['ms_in_capture_under_ncu']:8.3f} ms")
