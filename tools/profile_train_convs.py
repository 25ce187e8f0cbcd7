"""Which convolution shapes of one eager training step cost what on the GPU: torch.profiler with shapes, grouped by
(op, input shapes); CUDA time per group.  Evidence for which layers deserve own kernels."""
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
for _ in range(2):
    tr.train_step(batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    tr.train_step(batch)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True):
    dt = getattr(e, "self_device_time_total", None)
    if dt is None:
        dt = getattr(e, "self_cuda_time_total", 0)
    if dt > 0:
        rows.append((dt, e.count, e.key, str(e.input_shapes)[:150]))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"total self device time {tot / 1e3:.2f} ms")
print("-- ops (self device time, grouped by input shapes)")
for dt, n, k, shp in [r for r in rows if r[3] != "[]"][:60]:
    print(f"{dt / 1e3:8.3f} ms {n:4d}x  {k[:44]:44s} {shp}")
print("-- kernels")
for dt, n, k, shp in [r for r in rows if r[3] == "[]"][:45]:
    print(f"{dt / 1e3:8.3f} ms {n:4d}x  {k[:110]}")
