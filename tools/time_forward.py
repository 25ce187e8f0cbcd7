"""Per-launch CUDA-event timeline of one fused forward (shapes annotated).  usage: time_forward.py [pairs]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
h1, h2 = bench.make_inputs(0, P, 8)
from pwclonet_pylidarslam_b200.pwclonet import PWCLONet  # noqa: E402

dev = torch.device("cuda:0")
net = PWCLONet({"device": "cuda:0"})
net.load_state_dict({k: torch.from_numpy(v) for k, v in bench.make_weights().items()})
net = net.to(dev).eval()
d1, d2 = torch.from_numpy(h1).to(dev), torch.from_numpy(h2).to(dev)
eng = net.fused_engine()
with torch.no_grad():
    for _ in range(3):
        net(d1, None, d2, None)
    eng.verbose_timeline = True
    eng.timeline = []
    net(d1, None, d2, None)
torch.cuda.synchronize()
tot = 0.0
for name, s, e, _w in eng.timeline:
    ms = s.elapsed_time(e)
    tot += ms
    print(f"{ms:8.3f} ms  {name}")
print(f"{tot:8.3f} ms  total")
