"""Kernel timeline (CUPTI via torch.profiler) of graph-replayed training steps: which kernels run when, on which stream.
Writes a compact per-kernel table (start us, duration us, stream, name) for two steps: gpurun_out/train_graph_timeline.tsv"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
from pwclonet_pylidarslam_b200 import training as T  # noqa: E402

P, NP = 8, 16384
dev = torch.device("cuda:0")
h1, h2 = bench.make_inputs(50, P, P, n_points=NP)
torch.manual_seed(0)
tr = T.PWCLONetTrainer(T.PWCLONetTrainerConfig(num_points=NP, device="cuda:0", batch_size=P))
rng = np.random.default_rng(7)
q = rng.standard_normal((P, 4)).astype(np.float32) * 0.02 + np.array([1, 0, 0, 0], np.float32)
q /= np.linalg.norm(q, axis=-1, keepdims=True)
t = (rng.standard_normal((P, 3)) * np.array([0.05, 0.02, 0.3]) + np.array([0, 0, 1.0])).astype(np.float32)
batch = [torch.from_numpy(np.ascontiguousarray(h1.transpose(0, 2, 1))).to(dev), torch.from_numpy(np.ascontiguousarray(h2.transpose(0, 2, 1))).to(dev),
         torch.from_numpy(q).to(dev), torch.from_numpy(t).to(dev)]
tr.capture(batch)
for _ in range(5):
    tr.train_step_graphed(batch, next_batch=batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        tr.train_step_graphed(batch, next_batch=batch)
    torch.cuda.synchronize()
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/train_graph_trace.json"
prof.export_chrome_trace(out)
ev = [e for e in json.load(open(out))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
with open(out.replace(".json", ".tsv"), "w") as f:
    for e in ev:
        f.write(f"{e['ts'] - t0:.1f}\t{e['dur']:.1f}\t{e['args'].get('stream')}\t{e['name'][:100]}\n")
os.remove(out)
print(len(ev), "device activities", (ev[-1]["ts"] + ev[-1]["dur"] - t0) / 1e3, "ms")
