"""FPS with few clouds: the 8-CTA cluster kernel against one CTA per cloud (PWCLO_FPS_CLUSTER=0), bit-exact check
included.  usage: python tools/bench_fps_cluster.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pwclonet_pylidarslam_b200 import _ext, synthetic as syn  # noqa: E402

out = []
for B, N, m in ((16, 16384, 2048), (2, 8192, 2048), (16, 8192, 2048), (2, 16384, 2048), (8, 4096, 1024)):
    pts = [syn.make_pair(500 + i, N) for i in range((B + 1) // 2)]
    x = torch.from_numpy(np.stack([p[k] for p in pts for k in ("pc1", "pc2")][:B])).cuda().contiguous()
    res = {}
    for name, env in (("cluster", "2"), ("one_cta_per_cloud", "0")):
        os.environ["PWCLO_FPS_CLUSTER"] = env
        for _ in range(2):
            idx = _ext.furthest_point_sampling(x, m)
        ts = []
        for _ in range(5):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            idx = _ext.furthest_point_sampling(x, m)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        res[name] = (float(np.median(ts)), idx.cpu())
    same = bool(torch.equal(res["cluster"][1], res["one_cta_per_cloud"][1]))
    row = {"clouds": B, "points": N, "samples": m, "cluster_ms": res["cluster"][0], "one_cta_per_cloud_ms": res["one_cta_per_cloud"][0],
           "us_per_round_cluster": res["cluster"][0] * 1e3 / (m - 1), "identical_indices": same}
    out.append(row)
    print(json.dumps(row))
os.environ.pop("PWCLO_FPS_CLUSTER", None)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/fps_cluster_bench.json", "w"), indent=1)
