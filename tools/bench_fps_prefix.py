"""FPS prefix shortcut micro-benchmark: plain sampling vs sampling with the tie flag vs the chained levels (B200).
usage: python tools/bench_fps_prefix.py [clouds]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from pwclonet_pylidarslam_b200 import _ext  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
h1, h2 = bench.make_inputs(0, B // 2, 16)
xyz = torch.from_numpy(np.concatenate((h1, h2)).transpose(0, 2, 1).copy()).to(dev)


def t(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


out = {"clouds": B}
out["l1_plain_ms"] = t(lambda: _ext.furthest_point_sampling(xyz, 2048))
out["l1_tie_ms"] = t(lambda: _ext.furthest_point_sampling(xyz, 2048, return_tie=True))
idx, tie = _ext.furthest_point_sampling(xyz, 2048, return_tie=True)
out["l1_ties_set"] = int(tie.sum())
l1 = torch.gather(xyz, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
out["l2_plain_ms"] = t(lambda: _ext.furthest_point_sampling(l1, 1024))
out["l2_prefix_ms"] = t(lambda: _ext.furthest_point_sampling(l1, 1024, tie_in=tie, return_tie=True))
one = torch.ones_like(tie)
out["l2_forced_real_tie_ms"] = t(lambda: _ext.furthest_point_sampling(l1, 1024, tie_in=one, return_tie=True))
i2, t2 = _ext.furthest_point_sampling(l1, 1024, tie_in=tie, return_tie=True)
out["l2_is_arange"] = bool((i2 == torch.arange(1024, device=dev, dtype=torch.int32)).all())
out["l2_equals_plain"] = bool(torch.equal(i2, _ext.furthest_point_sampling(l1, 1024)))
print(json.dumps(out))
