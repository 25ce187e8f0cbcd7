"""FPS level-1 timing across kernel variants (env toggles are read per call)."""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pwclonet_pylidarslam_b200 import _ext  # noqa: E402
import bench  # noqa: E402

h1, h2 = bench.make_inputs(0, 64, 8)
x = torch.cat([torch.from_numpy(h1), torch.from_numpy(h2)]).cuda().permute(0, 2, 1).contiguous()
variants = {"plain": {"PWCLO_FPS_NO_SLAB": "1"}, "slab16x512": {}, "slab8x1024": {"PWCLO_FPS_SLAB8": "1"},
            "slab16 skip-all (wrong result, latency floor)": {"PWCLO_FPS_DBG_SKIPALL": "1"}}
for m in (1024, 2048):
    for name, env in variants.items():
        for k in ("PWCLO_FPS_NO_SLAB", "PWCLO_FPS_SLAB8", "PWCLO_FPS_DBG_SKIPALL"):
            os.environ.pop(k, None)
        os.environ.update(env)
        for _ in range(3):
            ref = _ext.furthest_point_sampling(x, m)
        ts = []
        for _ in range(10):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            i = _ext.furthest_point_sampling(x, m)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        print(f"m={m} {name:12s} min {min(ts):.3f} median {np.median(ts):.3f} max {max(ts):.3f} ms  checksum {int(i.sum())}")
