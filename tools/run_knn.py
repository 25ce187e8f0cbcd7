"""kNN call shapes of the forward, timed individually (presort + search), for ncu captures.
usage: run_knn.py B S N K [self]"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pwclonet_pylidarslam_b200 import _ext  # noqa: E402

B, S, N, K = [int(v) for v in sys.argv[1:5]]
dev = torch.device("cuda:0")
h1, h2 = bench.make_inputs(0, B, min(B, 8))
x1 = torch.from_numpy(h1).to(dev).permute(0, 2, 1).contiguous()
x2 = torch.from_numpy(h2).to(dev).permute(0, 2, 1).contiguous()
# emulate pyramid levels by FPS prefixes
i1 = _ext.furthest_point_sampling(x1, max(S, N)).long()
i2 = _ext.furthest_point_sampling(x2, max(S, N)).long()
g = lambda x, i, n: torch.gather(x, 1, i[:, :n, None].expand(-1, -1, 3)).contiguous()
q = g(x1, i1, S)
r = q if "self" in sys.argv else g(x2, i2, N)
if "rot" in sys.argv:      # a garbage pose estimate: the query cloud is rotated away from the reference cloud
    ang = 1.0
    R = torch.tensor([[1, 0, 0], [0, float(torch.cos(torch.tensor(ang))), -float(torch.sin(torch.tensor(ang)))],
                      [0, float(torch.sin(torch.tensor(ang))), float(torch.cos(torch.tensor(ang)))]], device=dev)
    q = (q @ R.T).contiguous() + torch.tensor([3.0, -2.0, 5.0], device=dev)
if "brute" in sys.argv:
    _ext.KNN_SORTED = False
for _ in range(3):
    _ext.knn(r, q, K)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    _ext.knn(r, q, K)
b.record()
torch.cuda.synchronize()
print(f"knn B{B} S{S} N{N} K{K}: {a.elapsed_time(b) / 10:.4f} ms")
