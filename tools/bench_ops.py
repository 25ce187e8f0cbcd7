"""Micro-benchmark of the stand-alone operators at BASELINE config-2 shapes (and at the batch the
full-forward bench uses).  CUDA-event timing on the launching stream, L2 flushed between reps.
Writes gpurun_out/ops_bench.json.  Optionally times the reference's own kernels (oracle/_ref)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pwclonet_pylidarslam_b200 import _ext  # noqa: E402

dev = torch.device("cuda:0")
PEAK = 6551.7
try:
    PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(min(ts))


def main():
    ref = None
    if "--ref" in sys.argv:
        from oracle import build_ref_ext
        ref = build_ref_ext.load_module()
    res = []
    g = torch.Generator(device=dev).manual_seed(0)
    for B in (8, 16, 128):
        xyz = (torch.randn(B, 8192, 3, device=dev, generator=g) * torch.tensor([20., 1., 20.], device=dev)).contiguous()
        levels = [(8192, 2048, 32, 3), (2048, 1024, 32, 16), (1024, 256, 16, 32)]
        cur = xyz
        for (N, M, K, C) in levels:
            idx = _ext.furthest_point_sampling(cur, M)
            med, mn = timeit(lambda: _ext.furthest_point_sampling(cur, M))
            alg = 4 * B * (3 * N + M)
            res.append(dict(op="fps", B=B, N=N, M=M, ms=med, ms_min=mn, alg_bytes=alg, gbs=alg / med / 1e6,
                            rounds_per_s=(M - 1) / (med * 1e-3)))
            if ref is not None:
                rmed, _ = timeit(lambda: ref.furthest_point_sampling(cur, M))
                res[-1]["ref_ms"] = rmed
            flipped = cur.transpose(1, 2).contiguous()
            new = _ext.gather_points(flipped, idx)
            med, mn = timeit(lambda: _ext.gather_points(flipped, idx))
            alg = 4 * B * (M + 3 * M + 3 * N)
            res.append(dict(op="gather", B=B, N=N, M=M, ms=med, ms_min=mn, alg_bytes=alg, gbs=alg / med / 1e6))
            new_xyz = new.transpose(1, 2).contiguous()
            nidx = _ext.knn(cur, new_xyz, K)
            med, mn = timeit(lambda: _ext.knn(cur, new_xyz, K))
            alg = 4 * B * (3 * M + 3 * N + M * K)
            res.append(dict(op="knn", B=B, N=N, S=M, K=K, ms=med, ms_min=mn, alg_bytes=alg, gbs=alg / med / 1e6,
                            gpairs_per_s=B * M * N / med / 1e6))
            for Cg in sorted({3, C}):
                feats = torch.randn(B, Cg, N, device=dev, generator=g)
                med, mn = timeit(lambda: _ext.group_points(feats, nidx))
                alg = 4 * B * (M * K + Cg * M * K + Cg * N)
                res.append(dict(op="group", B=B, C=Cg, N=N, S=M, K=K, ms=med, ms_min=mn, alg_bytes=alg,
                                gbs=alg / med / 1e6, frac_hbm=alg / med / 1e6 / PEAK))
                if ref is not None:
                    rmed, _ = timeit(lambda: ref.group_points(feats, nidx))
                    res[-1]["ref_ms"] = rmed
            cur = new_xyz
    # a launch large enough to be HBM-bound: group C=64 rows of N=2048 into S=2048,K=32 for 64 clouds
    B, C, N, S, K = 64, 64, 2048, 2048, 32
    feats = torch.randn(B, C, N, device=dev, generator=g)
    idx = torch.randint(0, N, (B, S, K), device=dev, dtype=torch.int32, generator=g)
    med, mn = timeit(lambda: _ext.group_points(feats, idx))
    alg = 4 * B * (S * K + C * S * K + C * N)
    res.append(dict(op="group_big", B=B, C=C, N=N, S=S, K=K, ms=med, ms_min=mn, alg_bytes=alg, gbs=alg / med / 1e6,
                    frac_hbm=alg / med / 1e6 / PEAK))
    if ref is not None:
        rmed, _ = timeit(lambda: ref.group_points(feats, idx))
        res[-1]["ref_ms"] = rmed
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/ops_bench.json", "w"), indent=1)
    for r in res:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
