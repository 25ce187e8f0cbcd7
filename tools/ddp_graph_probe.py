"""2-rank check of the data-parallel training step replayed as one CUDA graph (NCCL all-reduce captured inside):
graphed losses == eager losses of an identically initialised trainer, then a clean teardown (trainer.close()).
Launched by tests/test_training_gpu.py::test_ddp_graphed_step_equals_eager_step and tools/try_ddp_graph.sh under
torchrun; prints `rank r: OK` per rank.  Dropout is switched off (eager and replay draw different masks)."""
import os
import sys
import threading

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F_

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pwclonet_pylidarslam_b200 import synthetic as syn, training as T  # noqa: E402

F_.dropout = lambda x, p=0.5, training=True, inplace=False: x
threading.Timer(float(os.environ.get("PROBE_LIMIT_S", "140")), lambda: os._exit(3)).start()   # never hang a GPU box
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
x1, x2, gt = syn.make_batch(900 + 10 * local, 2, 4096)
batch = [torch.from_numpy(np.ascontiguousarray(x1.transpose(0, 2, 1))).to(dev), torch.from_numpy(np.ascontiguousarray(x2.transpose(0, 2, 1))).to(dev),
         torch.from_numpy(gt[:, 3:]).to(dev), torch.from_numpy(gt[:, :3]).to(dev)]
batch2 = [b.flip(0).contiguous() for b in batch]
# a small learning rate keeps the comparison about the step, not about how fast two trajectories diverge under the
# run-to-run noise of the scatter-add atomics
cfg = T.PWCLONetTrainerConfig(num_points=4096, device=str(dev), optimizer_learning_rate=1e-5)
torch.manual_seed(0)
a = T.PWCLONetTrainer(cfg)
b = T.PWCLONetTrainer(cfg)
b.prediction_module_.load_state_dict(a.prediction_module_.state_dict())
def sync_b_to_a():
    b.prediction_module_.load_state_dict(a.prediction_module_.state_dict())
    b.loss_module_.load_state_dict(a.loss_module_.state_dict())
    b._optimizer.load_state_dict(a._optimizer.state_dict())


for _ in range(2):
    a.train_step(batch)
b.capture(batch, warmup=1)
# Every comparison starts from IDENTICAL state (a -> b), so it checks one step, not how fast two trajectories of a chaotic
# system (random-init network, 2-pair BatchNorm, scatter-add atomics) drift apart: the loss (the forward) must agree to
# 1e-5, the updated parameters to a few learning rates (Adam's normalised update can flip on gradient entries that are
# pure atomics noise).
lr = cfg.optimizer_learning_rate
report = []
for i in range(4):
    bt = batch if i % 2 == 0 else batch2
    sync_b_to_a()
    la = float(a.train_step(bt)[0])
    lb = float(b.train_step_graphed(bt, next_batch=(batch2 if i % 2 == 0 else batch) if i < 2 else None)[0])
    dp = (a.arena.param - b.arena.param).abs()
    report.append((la, lb, float(dp.max()) / lr, float(dp.mean()) / lr))
    assert abs(la - lb) <= 1e-5 * abs(la), (i, la, lb)
    assert float(dp.max()) <= 3.0 * lr and float(dp.mean()) <= 0.2 * lr, (i, float(dp.max()) / lr, float(dp.mean()) / lr)
print(f"rank {local}: (eager loss, graphed loss, max |dparam| / lr, mean |dparam| / lr) per step: {report}", flush=True)
# both ranks hold the same parameters after data-parallel steps
p = b.arena.param.clone()
dist.all_reduce(p, op=dist.ReduceOp.MAX)
assert torch.equal(p, b.arena.param), "ranks diverged"
b.close()
a.close()
dist.destroy_process_group()
print(f"rank {local}: OK", flush=True)
os._exit(0)
