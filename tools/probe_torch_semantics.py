"""One-off GPU probe: pins the two implementation-defined behaviours of torch
that the reference's knn_point inherits on CUDA.  Writes gpurun_out/probe_torch.json."""
import json
import os

import numpy as np
import torch

out = {}
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
x = torch.from_numpy(rng.standard_normal((1, 64, 4096, 3)).astype(np.float32) * 7).to(dev)
sq = x ** 2
s = torch.sum(sq, dim=-1)
a, b, c = sq[..., 0], sq[..., 1], sq[..., 2]
out["sum3_equals_(x+y)+z"] = float(((a + b) + c == s).float().mean())
out["sum3_equals_(x+z)+y"] = float(((a + c) + b == s).float().mean())
out["sum3_equals_x+(y+z)"] = float((a + (b + c) == s).float().mean())
# same through the exact knn_point expression (repeat / sub / pow / sum on a [B,S,N,3] tensor)
xyz = torch.from_numpy(rng.standard_normal((2, 2048, 3)).astype(np.float32) * 5).to(dev)
q = xyz[:, :512].contiguous()
diff = q.unsqueeze(2).repeat(1, 1, 2048, 1) - xyz.unsqueeze(1).repeat(1, 512, 1, 1)
s2 = torch.sum(diff ** 2, dim=-1)
d = diff * diff
out["knn_sum_equals_(x+y)+z"] = float((((d[..., 0] + d[..., 1]) + d[..., 2]) == s2).float().mean())
out["knn_sum_equals_(x+z)+y"] = float((((d[..., 0] + d[..., 2]) + d[..., 1]) == s2).float().mean())
# pow(x,2) == x*x ?
out["pow2_equals_mul"] = bool(torch.equal(diff ** 2, diff * diff))
# topk tie order
t = torch.tensor([[5., 1., 3., 1., 1., 0., 1., 9., 1.]], device=dev)
v, i = torch.topk(t, 4, largest=False)
out["topk_ties_k4"] = i.tolist()
big = torch.ones(1, 5000, device=dev)
big[0, 1234] = 0.5
v, i = torch.topk(big, 8, largest=False)
out["topk_ties_big_k8"] = i.tolist()
r = torch.from_numpy(rng.integers(0, 50, size=(4, 3000)).astype(np.float32)).to(dev)
v, i = torch.topk(r, 32, largest=False)
asc = []
for row_v, row_i in zip(v.cpu().numpy(), i.cpu().numpy()):
    ok = all(row_i[j] < row_i[j + 1] for j in range(31) if row_v[j] == row_v[j + 1])
    asc.append(bool(ok))
out["topk_equal_values_index_ascending_k32"] = asc
out["torch"] = torch.__version__
out["device"] = torch.cuda.get_device_name(0)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe_torch.json", "w"), indent=1)
print(json.dumps(out, indent=1))
