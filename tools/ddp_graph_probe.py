"""2-rank probe for tools/try_ddp_graph.sh: capture the data-parallel training step (PWCLO_GRAPH_DDP=1) on a small problem
and compare three replays with the eager step of an identically initialised trainer."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pwclonet_pylidarslam_b200 import synthetic as syn, training as T  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
x1, x2, gt = syn.make_batch(900 + 10 * local, 2, 4096)
batch = [torch.from_numpy(np.ascontiguousarray(x1.transpose(0, 2, 1))).to(dev), torch.from_numpy(np.ascontiguousarray(x2.transpose(0, 2, 1))).to(dev),
         torch.from_numpy(gt[:, 3:]).to(dev), torch.from_numpy(gt[:, :3]).to(dev)]
torch.manual_seed(0)
tr = T.PWCLONetTrainer(T.PWCLONetTrainerConfig(num_points=4096, device=str(dev)))
print(f"rank {local}: eager step", float(tr.train_step(batch)[0]), flush=True)
tr.capture(batch, warmup=1)
for i in range(3):
    print(f"rank {local}: replay {i}", float(tr.train_step_graphed(batch)[0]), flush=True)
dist.barrier()
dist.destroy_process_group()
print(f"rank {local}: done", flush=True)
