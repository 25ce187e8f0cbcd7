"""Times the fused cost-volume kernels alone on random inputs (level-1 shapes by default).
usage: bench_layer.py [B S N K C]"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pwclonet_pylidarslam_b200.pwclonet import PWCLONet  # noqa: E402

B, S, N, K, C = [int(v) for v in sys.argv[1:6]] if len(sys.argv) >= 6 else (64, 2048, 2048, 6, 16)
dev = torch.device("cuda:0")
net = PWCLONet({"device": "cuda:0"})
net.load_state_dict({k: torch.from_numpy(v) for k, v in bench.make_weights().items()})
eng = net.to(dev).eval().fused_engine()
g = torch.Generator(device=dev).manual_seed(0)
prefix = {64: "cost_volume" if K == 32 else "pose_warp_refinement_3.cost_volume", 32: "pose_warp_refinement_2.cost_volume",
          16: "pose_warp_refinement_1.cost_volume"}[C]
wxyz = torch.randn(B, S, 3, device=dev, generator=g)
xyz2 = torch.randn(B, N, 3, device=dev, generator=g)
f1 = torch.randn(B, S, C, device=dev, generator=g)
f2 = torch.randn(B, N, C, device=dev, generator=g)
idx_q = torch.randint(0, N, (B, S, K), device=dev, dtype=torch.int32, generator=g)
idx_s = torch.randint(0, S, (B, S, 4), device=dev, dtype=torch.int32, generator=g)
for _ in range(3):
    eng.cost_volume(prefix, wxyz, f1, xyz2, f2, idx_q, idx_s)
torch.cuda.synchronize()
eng.timeline = []
for _ in range(5):
    eng.cost_volume(prefix, wxyz, f1, xyz2, f2, idx_q, idx_s)
torch.cuda.synchronize()
acc = {}
for name, s, e, _w in eng.timeline:
    acc.setdefault(name, []).append(s.elapsed_time(e))
for k, v in acc.items():
    print(f"{k}: {sorted(v)[len(v) // 2]:.4f} ms  (B{B} S{S} N{N} K{K} C{C}, debug={os.environ.get('PWCLO_TC_DEBUG', '0')})")
