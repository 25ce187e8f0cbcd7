"""group_points at the HBM-bound size (C=64, N=2048, S=2048, K=32, B=64: 1.07 GB written) under different timing
protocols, next to a write-only fill and a copy of the same size."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pwclonet_pylidarslam_b200 import _ext  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
B, C, N, S, K = 64, 64, 2048, 2048, 32
feats = torch.randn(B, C, N, device=dev, generator=g)
idx = torch.randint(0, N, (B, S, K), device=dev, dtype=torch.int32, generator=g)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out_bytes = 4 * B * C * S * K
alg = 4 * B * (S * K + C * S * K + C * N)


def ev(fn, n, pre=None):
    ts = []
    for _ in range(n):
        if pre:
            pre()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]


for _ in range(3):
    _ext.group_points(feats, idx)
print("group, flush + sync before each :", alg / ev(lambda: _ext.group_points(feats, idx), 7, lambda: (flush.zero_(), torch.cuda.synchronize())) / 1e6, "GB/s")
print("group, flush enqueued right before:", alg / ev(lambda: _ext.group_points(feats, idx), 7, lambda: flush.zero_()) / 1e6, "GB/s")
print("group, back to back               :", alg / ev(lambda: _ext.group_points(feats, idx), 7) / 1e6, "GB/s")
buf = torch.empty(out_bytes // 4, device=dev)
src = torch.empty(out_bytes // 4, device=dev)
print("fill_ (write only), same bytes    :", out_bytes / ev(lambda: buf.fill_(1.0), 7, lambda: (flush.zero_(), torch.cuda.synchronize())) / 1e6, "GB/s written")
print("copy_ (read + write)              :", 2 * out_bytes / ev(lambda: buf.copy_(src), 7, lambda: (flush.zero_(), torch.cuda.synchronize())) / 1e6, "GB/s r+w")
