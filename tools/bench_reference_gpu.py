"""The like-for-like GPU baseline: the UNMODIFIED reference PWCLONet on the B200 --
  ref      reference python + the reference's own CUDA extension recompiled for sm_100a (oracle/_ref)
  ref+ext  reference python + this repository's drop-in extension (register_as_pointnet2_ops_ext)
  ref+ext+knn  ... and this repository's knn_point instead of the materialising torch one
  fused    this repository's fused forward (CUDA graph)
at B = 1 / 8 / 64 frame pairs of 8192 points, fp32 (TF32 disabled), eval mode, no_grad, CUDA-event timed after warm-up.
Writes gpurun_out/reference_gpu_bench.json.     usage: python tools/bench_reference_gpu.py [Bs comma list]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from oracle import build_ref_ext, ref_shim  # noqa: E402
from pwclonet_pylidarslam_b200 import _ext  # noqa: E402
from pwclonet_pylidarslam_b200.pytorch_utils import knn_point  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
Bs = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 8, 64]
ref_ext = build_ref_ext.load_module()
ref = ref_shim.load_reference(ext_module=ref_ext, device="cuda:0").to(dev).eval()
mods = ref_shim.reference_modules()
p2u, ptu = mods["pointnet2_utils"], mods["pytorch_utils"]
w = bench.make_weights()
sd = {k: torch.from_numpy(v) for k, v in w.items()}
ref.load_state_dict(sd)
ours = bench.build_net(dev, w)
h1, h2 = bench.make_inputs(0, max(Bs), 16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
torch_knn = ptu.knn_point


def timeit(fn, reps):
    with torch.no_grad():
        fn()
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
    return float(np.median(ts))


rows = []
for B in Bs:
    a, b = torch.from_numpy(h1[:B]).to(dev), torch.from_numpy(h2[:B]).to(dev)
    row = {"pairs": B}
    variants = (("ref", ref_ext, torch_knn), ("ref+ext", _ext, torch_knn), ("ref+ext+knn", _ext, knn_point))
    for name, ext, knn in variants:
        p2u._ext, ptu.knn_point = ext, knn
        try:
            torch.cuda.reset_peak_memory_stats()
            ms = timeit(lambda: ref(a, None, b, None), 3 if B >= 32 else 5)
            row[name] = {"ms": ms, "pairs_per_s": B / ms * 1e3, "peak_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
        except Exception as e:      # e.g. out of memory in the materialised [B,S,N,3] kNN tensors
            row[name] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
            torch.cuda.empty_cache()
    p2u._ext, ptu.knn_point = ref_ext, torch_knn
    ms = timeit(lambda: ours(a, None, b, None), 10)
    row["fused"] = {"ms": ms, "pairs_per_s": B / ms * 1e3}
    if "ms" in row["ref"]:
        row["speedup_vs_ref"] = row["ref"]["ms"] / ms
    rows.append(row)
    print(json.dumps(row), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"what": "unmodified reference PWCLONet on the B200 vs the fused forward, fp32, 8192 points", "rows": rows},
          open(os.path.join(ROOT, "gpurun_out", "reference_gpu_bench.json"), "w"), indent=1)
