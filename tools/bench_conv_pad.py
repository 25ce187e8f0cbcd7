"""1x1 convolutions of the training step whose input-channel count is not a multiple of 4 (19, 67, 42, 138, 10, 74, 35):
library forward + backward time as is against zero-padded channels (multiple of 4 / 8)."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dev = torch.device("cuda:0")
SHAPES = [(19, 16, 1024, 32), (67, 128, 2048, 8), (67, 128, 1024, 8), (42, 128, 2048, 6), (138, 128, 256, 32), (10, 64, 2048, 6),
          (10, 64, 2048, 4), (10, 64, 256, 32), (74, 128, 1024, 6), (35, 32, 256, 16), (67, 128, 256, 8), (10, 64, 1024, 6)]


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


for ci, co, S, K in SHAPES:
    out = []
    for mult in (1, 4, 8):
        cp = (ci + mult - 1) // mult * mult
        x = torch.randn(8, cp, S, K, device=dev, requires_grad=True)
        w = torch.randn(co, cp, 1, 1, device=dev, requires_grad=True)
        dy = torch.randn(8, co, S, K, device=dev)

        def fwd():
            return F.conv2d(x, w)

        def both():
            y = F.conv2d(x, w)
            torch.autograd.grad(y, (x, w), dy)

        tf = timed(fwd)
        tb = timed(both) - tf
        out.append(f"ci={cp:3d}: fwd {tf * 1e3:6.1f} us  bwd {tb * 1e3:6.1f} us")
    print(f"[8,{ci}->{co},{S},{K}]  " + "  |  ".join(out), flush=True)
