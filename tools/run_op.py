"""Runs one operator a few times at a fixed shape (for ncu captures).  usage: run_op.py knn|fps|group [B]"""
import sys
import os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pwclonet_pylidarslam_b200 import _ext  # noqa: E402

op = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
xyz = (torch.randn(B, 8192, 3, device=dev, generator=g) * torch.tensor([20., 1., 20.], device=dev)).contiguous()
if op == "fps":
    for _ in range(3):
        _ext.furthest_point_sampling(xyz, 2048)
elif op == "knn":
    q = xyz[:, :2048].contiguous()
    for _ in range(3):
        _ext.knn(xyz, q, 32)
elif op == "group":
    feats = torch.randn(B, 16, 2048, device=dev, generator=g)
    idx = torch.randint(0, 2048, (B, 1024, 32), device=dev, dtype=torch.int32, generator=g)
    for _ in range(3):
        _ext.group_points(feats, idx)
torch.cuda.synchronize()
print("ok")
