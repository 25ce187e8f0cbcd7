"""The program ncu profiles in round 2: three warm-up forwards, then ONE launch-by-launch forward of 64 frame pairs between
cudaProfilerStart / Stop (ncu --profile-from-start off), no CUDA graph, one stream -- the same kernels, in the same order
and on the same inputs, as a bench.py step.   usage: python tools/ncu_forward.py [pairs]"""
import os
import sys

os.environ["PWCLO_OVERLAP"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
h1, h2 = bench.make_inputs(0, P, 16)
dev = torch.device("cuda:0")
net = bench.build_net(dev, bench.make_weights())
eng = net.fused_engine()
d1, d2 = torch.from_numpy(h1).to(dev), torch.from_numpy(h2).to(dev)
with torch.no_grad():
    for _ in range(3):
        eng.forward(d1, d2)
    torch.cuda.synchronize()
    n0 = eng.launches
    torch.cuda.profiler.start()
    eng.forward(d1, d2)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("C-ABI calls in the profiled forward:", eng.launches - n0)
