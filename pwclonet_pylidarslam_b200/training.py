"""Training-side rows of SURVEY 8 (F15, N1): the PWCLO-Net loss, the optimiser, the data-parallel
gradient exchange and a minimal trainer with reference-compatible checkpoints.

Mirrors, with the reference's names and argument meaning:
  * `ExponentialWeights`, `PWCLONetLossConfig`, `_PWCLONetLossModule`  slam/training/loss_modules.py:147-196, 303-544
  * `_PWCLONetPredictionModule`                                       slam/training/prediction_modules.py:99-153
  * `PWCLONetTrainer` (train_epoch / save_checkpoint / load_checkpoint, Adam + cosine or clipped
    exponential LR schedule + BN momentum schedule)                   slam/training/trainer.py:309-323, 546-676, 840-907;
                                                                      train.py:146-180, 309-323

B200-first differences (results equal, execution not):
  * the loss and its whole gradient are ONE kernel launch (`pwclo_pose_loss`) instead of ~120
    element-wise launches plus their autograd replay;
  * every trainable tensor of the prediction and loss modules is a view into ONE flat fp32 arena,
    and so is its `.grad`: data-parallel training needs exactly one NCCL all-reduce of 775 070
    floats per step over NVLink and the Adam update is one launch (`pwclo_adam_step`,
    the 1/world_size averaging folded into it);
  * no host synchronisation inside a step (the reference calls `.cpu()` on loss and predictions
    every iteration, trainer.py:626-633; here logs stay on the device until they are read).
There is no CPU path: the loss and the optimiser raise on CPU tensors.
"""
import ctypes
import math
import os
from dataclasses import dataclass, field
from typing import List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn
from torch.autograd import Function

from . import _lib
from .pytorch_utils import BNMomentumScheduler


# ---------------------------------------------------------------------------------------------- loss
@dataclass
class PWCLONetLossConfig:
    """slam/training/loss_modules.py:303-324 (hydra-free)."""
    mode: str = "supervised"
    loss_degrees: bool = False
    loss_weights: List[float] = field(default_factory=lambda: [1.0, 1.0])
    with_exp_weights: bool = True
    init_weights: List[float] = field(default_factory=lambda: [0.0, -2.5])
    loss_option: str = "l2_norm"
    nb_levels: int = 4
    device: str = "cpu"
    scalar_last: bool = False


class ExponentialWeights(nn.Module):
    """loss = sum_k loss_k * exp(-s_k) + s_k with trained s (loss_modules.py:147-196)."""

    def __init__(self, num_losses: int, init_weights: list):
        super().__init__()
        assert len(init_weights) == num_losses
        self.s_param = nn.Parameter(torch.tensor(init_weights, dtype=torch.float32), requires_grad=True)
        self.num_losses = num_losses

    def forward(self, list_losses: list):
        assert len(list_losses) == self.num_losses
        loss, s_params = 0.0, []
        for i in range(self.num_losses):
            s = self.s_param[i]
            loss = loss + (list_losses[i] * torch.exp(-s) + s)
            s_params.append(s.detach())
        return loss, s_params


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _need_cuda_f32(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32):
        raise RuntimeError(f"{name} must be a float32 CUDA tensor (there is no CPU path)")


class PoseLossFunction(Function):
    """(pred[B,4,7], gt[B,7], s[2]) -> (loss, terms[16]); value and gradient from one launch."""

    @staticmethod
    def forward(ctx, pred, gt, s, with_exp):
        for t, n in ((pred, "pred_params"), (gt, "gt_params"), (s, "s_param")):
            _need_cuda_f32(t, n)
        pred, gt, s = pred.contiguous(), gt.contiguous(), s.contiguous()
        B = pred.shape[0]
        terms = torch.empty(16, dtype=torch.float32, device=pred.device)
        gp = torch.empty_like(pred)
        gs = torch.empty(2, dtype=torch.float32, device=pred.device)
        with torch.cuda.device(pred.device):
            _lib.check(_lib.lib().pwclo_pose_loss(_p(pred), _p(gt), _p(s), B, 1 if with_exp else 0, _p(terms), _p(gp),
                                                  _p(gs), _lib.stream_ptr()), "pose_loss")
        ctx.save_for_backward(gp, gs)
        ctx.mark_non_differentiable(terms)
        return terms[0], terms

    @staticmethod
    def backward(ctx, g_loss, _g_terms):
        gp, gs = ctx.saved_tensors
        return g_loss * gp, None, g_loss * gs, None


class _PWCLONetLossModule(nn.Module):
    """Supervised loss of PWCLO-Net (loss_modules.py:329-544): per level the L2 norm of the
    normalised-quaternion error (mean over the batch) and mean sqrt((t - t_gt)^2 + 1e-10) over batch
    and axes, combined by ExponentialWeights (or fixed weights) and summed with level weights
    1.6 / 0.8 / 0.4 / 0.2 (coarsest .. finest).  forward(pred_params[B,4,7], gt_params[B,7]) ->
    (loss, log_dict) with the reference's log keys; log values are 0-dim device tensors."""

    def __init__(self, config, pose=None):
        super().__init__()
        if isinstance(config, dict):
            config = PWCLONetLossConfig(**{k: v for k, v in config.items() if k in PWCLONetLossConfig.__dataclass_fields__})
        self.config = config
        self.pose = pose
        self.exp_weighting: Optional[ExponentialWeights] = None
        self.weights: Optional[list] = None
        self.degrees = config.loss_degrees
        if config.with_exp_weights:
            self.exp_weighting = ExponentialWeights(2, list(config.init_weights))
        else:
            self.weights = list(config.loss_weights)
            assert len(self.weights) == 2
            self.register_buffer("_fixed", torch.tensor(self.weights, dtype=torch.float32), persistent=False)
        assert config.loss_option in ["l1", "l2", "l2_norm"]
        self.loss_config = config.loss_option
        self.nb_levels = config.nb_levels

    def forward(self, pred_params, gt_params):
        assert pred_params.dim() == 3 and pred_params.size(1) == 4 and pred_params.size(2) == 7
        assert gt_params.size(1) == 7 and gt_params.size(0) == pred_params.size(0)
        with_exp = self.exp_weighting is not None
        s = self.exp_weighting.s_param if with_exp else self._fixed
        loss, t = PoseLossFunction.apply(pred_params, gt_params, s, with_exp)
        log = {}
        for l in range(4):
            log[f"loss_rot_l{l + 1}"] = t[5 + l]
            log[f"loss_trans_l{l + 1}"] = t[9 + l]
        if with_exp:
            for l in range(4):
                log[f"s_rot_l{l + 1}"] = t[14]
                log[f"s_trans_l{l + 1}"] = t[13]
        for l in range(4):
            log[f"loss_l{l + 1}"] = t[1 + l]
        log["loss"] = loss
        if with_exp:
            log["s_param_trans"] = t[13]
            log["s_param_rot"] = t[14]
        return loss, log


# ------------------------------------------------------------------------------- prediction module
class _PWCLONetPredictionModule(nn.Module):
    """slam/training/prediction_modules.py:99-153: takes the collated batch (dict with `numpy_pc_0/1`
    or a list whose first two entries are [B,N,>=3] clouds), keeps xyz, hands [B,3,N] to PWCLONet."""

    PC_KEY = "numpy_pc"

    def __init__(self, config, pose=None):
        super().__init__()
        from .pwclonet import PWCLONet
        cfg = dict(config)
        self.config = cfg
        self.pose = pose
        self.device = torch.device(cfg.get("device", "cuda:0"))
        self.num_input_channels = cfg.get("num_input_channels", 3)
        self.sequence_len = cfg.get("sequence_len", 2)
        assert self.sequence_len == 2, "PWCLONet is developed to only accept 2 frames"
        self.num_points = cfg.get("num_points", 8192)
        self.nb_levels = cfg.get("nb_levels", 4)
        net_cfg = dict(cfg.get("posenet_config", {}) or {})
        net_cfg.update(sequence_len=self.sequence_len, num_input_channels=self.num_input_channels,
                       num_points=self.num_points, nb_levels=self.nb_levels, device=str(self.device),
                       scalar_last=cfg.get("scalar_last", False))
        self.pwclonet = PWCLONet(net_cfg)

    def _clouds(self, data):
        if isinstance(data, dict):
            pcs = []
            for i in range(self.sequence_len):
                key = f"{self.PC_KEY}_{i}"
                if key not in data:
                    raise RuntimeError(f"key `{key}` not found in data when running the prediction module")
                pcs.append(data[key])
        elif isinstance(data, (list, tuple)):
            pcs = [data[0], data[1]]
        else:
            raise RuntimeError("Input data should be either dict or list")
        xyz, feats = [], []
        for pc in pcs:
            xyz.append(pc[:, :self.num_points, :3].permute(0, 2, 1).contiguous())
            feats.append(pc[:, :self.num_points, 3:].permute(0, 2, 1).contiguous() if pc.size(-1) > 3 else None)
        return xyz, feats

    def forward(self, data, bn_decay=None, geoms=None):
        xyz, feats = self._clouds(data)
        return self.pwclonet(xyz[0], feats[0], xyz[1], feats[1], bn_decay=bn_decay, geoms=geoms)

    def geometry(self, data):
        """the coordinates-only part of the forward for `data` (PWCLONet.pyramid_geometry): valid for forward(data, geoms=...)"""
        xyz, _ = self._clouds(data)
        return self.pwclonet.pyramid_geometry(xyz[0], xyz[1])


# ------------------------------------------------------------------ flat arena, all-reduce, Adam
class FlatArena:
    """Every trainable tensor of `modules` becomes a view into one contiguous fp32 buffer (`param`),
    its gradient a view into `grad`.  Order = `module.parameters()` order of each module in turn,
    i.e. the order torch.optim.Adam's state dict numbers them (trainer.py:310-316)."""

    def __init__(self, modules):
        self.groups = [[p for p in m.parameters() if p.requires_grad] for m in modules]
        self.params = [p for g in self.groups for p in g]
        if not self.params:
            raise RuntimeError("no trainable parameters")
        dev = self.params[0].device
        # each tensor starts on a 16-byte boundary so that float4 access never straddles two tensors
        self.offsets, off = [], 0
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise RuntimeError("FlatArena needs float32 parameters on one device")
            self.offsets.append(off)
            off += (p.numel() + 3) // 4 * 4
        self.numel = off
        self.count = sum(p.numel() for p in self.params)
        self.param = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(off, dtype=torch.float32, device=dev)
        for p, o in zip(self.params, self.offsets):
            n = p.numel()
            self.param[o:o + n].copy_(p.data.reshape(-1))
            p.data = self.param[o:o + n].view(p.shape)
            p.grad = self.grad[o:o + n].view(p.shape)

    # -- one step without per-parameter accumulation kernels ---------------------------------------------------------
    # With `.grad` pre-set, autograd's AccumulateGrad adds into it: one small add kernel per parameter tensor (318 per
    # step, 6 % of the launches of a training step, profiles/round2_train_launches_summary.txt).  detach_grads() before
    # the backward lets autograd hand over its own buffers instead; collect_grads() then gathers them into the flat
    # arena with a handful of multi-tensor copies and restores the invariant "p.grad is a view into arena.grad".
    def detach_grads(self):
        for p in self.params:
            p.grad = None

    def collect_grads(self):
        views = self.views(self.grad)
        have = [(v, p.grad) for p, v in zip(self.params, views) if p.grad is not None]
        if len(have) != len(views):
            self.grad.zero_()                      # parameters the loss did not reach keep a zero gradient
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        for p, v in zip(self.params, views):
            p.grad = v

    def zero_grad(self):
        self.grad.zero_()
        for p, o in zip(self.params, self.offsets):   # re-attach if something replaced a .grad
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * o:
                p.grad = self.grad[o:o + p.numel()].view(p.shape)

    def views(self, flat):
        return [flat[o:o + p.numel()].view(p.shape) for p, o in zip(self.params, self.offsets)]


def all_reduce_gradients(arena: FlatArena):
    """The one collective of data-parallel training: sum of the flat gradient arena over
    ranks.  Returns the factor the optimiser must apply (1/world_size) instead of rescaling here."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(arena.grad, op=dist.ReduceOp.SUM)
        return 1.0 / dist.get_world_size()
    return 1.0


class FlatAdam:
    """torch.optim.Adam semantics (L2 weight decay in the gradient, bias correction) on a FlatArena; one
    `pwclo_adam_step_dev` call per step.  The step counter and the learning rate live in device memory, so a
    captured CUDA graph of the step stays valid across steps and learning-rate changes.
    state_dict()/load_state_dict() use torch.optim.Adam's layout so that checkpoints move between this trainer
    and the reference's (trainer.py:884)."""

    def __init__(self, arena: FlatArena, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if not arena.param.is_cuda:
            raise RuntimeError("FlatAdam needs CUDA parameters (there is no CPU path)")
        self.arena, self.betas, self.eps, self.weight_decay = arena, tuple(betas), float(eps), float(weight_decay)
        self.initial_lr = float(lr)
        self.exp_avg = torch.zeros_like(arena.param)
        self.exp_avg_sq = torch.zeros_like(arena.param)
        dev = arena.param.device
        self._state_i = torch.zeros(4, dtype=torch.int32, device=dev)       # [0] = completed steps
        self._state_f = torch.zeros(4, dtype=torch.float32, device=dev)     # [0] = lr, [1..2] = per-step scratch
        self._lr = None
        self.lr = float(lr)

    @property
    def lr(self):
        return self._lr

    @lr.setter
    def lr(self, value):
        self._lr = float(value)
        self._state_f[0:1].fill_(self._lr)          # stream-ordered: takes effect from the next (replayed) step

    @property
    def steps(self):
        return int(self._state_i[0])                 # device -> host read; not on the step path

    @steps.setter
    def steps(self, value):
        self._state_i[0:1].fill_(int(value))

    def zero_grad(self, set_to_none=False):
        self.arena.zero_grad()

    def step(self, grad_scale=1.0):
        a = self.arena
        with torch.cuda.device(a.param.device):
            _lib.check(_lib.lib().pwclo_adam_step_dev(_p(a.param), _p(a.grad), _p(self.exp_avg), _p(self.exp_avg_sq), a.numel,
                                                      _p(self._state_i), _p(self._state_f), self.betas[0], self.betas[1],
                                                      self.eps, self.weight_decay, float(grad_scale), _lib.stream_ptr()),
                       "adam_step")

    @property
    def param_groups(self):
        return [{"lr": self.lr, "initial_lr": self.initial_lr}]

    def state_dict(self):
        state, groups, k = {}, [], 0
        steps = self.steps
        m, v = self.arena.views(self.exp_avg), self.arena.views(self.exp_avg_sq)
        for g in self.arena.groups:
            ids = list(range(k, k + len(g)))
            for i in ids:
                if steps:
                    state[i] = {"step": torch.tensor(float(steps)), "exp_avg": m[i].clone(), "exp_avg_sq": v[i].clone()}
            groups.append({"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay,
                           "amsgrad": False, "maximize": False, "foreach": None, "capturable": False,
                           "differentiable": False, "fused": None, "initial_lr": self.initial_lr, "params": ids})
            k += len(g)
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd):
        m, v = self.arena.views(self.exp_avg), self.arena.views(self.exp_avg_sq)
        steps = 0
        for i, st in sd.get("state", {}).items():
            i = int(i)
            m[i].copy_(st["exp_avg"])
            v[i].copy_(st["exp_avg_sq"])
            steps = max(steps, int(float(st["step"])))
        self.steps = steps
        if sd.get("param_groups"):
            g = sd["param_groups"][0]
            self.lr = float(g.get("lr", self.lr))
            self.initial_lr = float(g.get("initial_lr", self.lr))
            self.betas = tuple(g.get("betas", self.betas))
            self.eps = float(g.get("eps", self.eps))
            self.weight_decay = float(g.get("weight_decay", self.weight_decay))


# ------------------------------------------------------------------------------------- schedules
def cosine_lr(base_lr, epoch, t_max, eta_min):
    """closed form of torch CosineAnnealingLR(T_max, eta_min) as train.py:309-314 builds it"""
    return eta_min + (base_lr - eta_min) * (1.0 + math.cos(math.pi * epoch / t_max)) / 2.0


def exponential_lr(base_lr, epoch, gamma, decay_clip):
    """PWCLONetEexponentialScheduler closed form (train.py:146-180): max(lr0 * gamma^epoch, clip)"""
    return max(base_lr * gamma ** epoch, decay_clip)


def bn_momentum(epoch, init=0.5, rate=0.5, step=4, mx=0.99):
    """train.py:318-322: min(1 - init * rate^(epoch // step), max)"""
    return min(1 - init * rate ** (int(epoch / step)), mx)


@dataclass
class PWCLONetTrainerConfig:
    """The fields of PWCLONetConfig / ATrainerConfig that drive the numerical path (train.py:195-220,
    trainer.py:97-115, config/train_pwclonet.yaml)."""
    num_epochs: int = 100
    batch_size: int = 8
    num_points: int = 8192
    device: str = "cuda:0"
    optimizer_type: str = "adam"
    optimizer_learning_rate: float = 0.001
    optimizer_beta: float = 0.9
    optimizer_momentum: float = 0.999
    optimizer_weight_decay: float = 0.001
    optimizer_scheduler_decay: float = 0.7
    scheduler_decay_clip: float = 0.000001
    coslr: bool = True
    bn_momentum_init: float = 0.5
    bn_decay_rate: float = 0.5
    bn_decay_step: int = 4
    bn_momentum_max: float = 0.99
    loss: PWCLONetLossConfig = field(default_factory=PWCLONetLossConfig)
    scalar_last: bool = False


class PWCLONetTrainer:
    """Minimal equivalent of ATrainer / PWCLONetTrainer without hydra / wandb / tqdm.

    `train_step(batch)` = zero_grad -> prediction -> loss -> backward -> (all-reduce) -> Adam, the body
    of ATrainer.train_epoch's loop (trainer.py:594-632); `batch` is the collated reference batch
    `[pc1[B,N,3], pc2[B,N,3], q_gt[B,4], t_gt[B,3], ...]` (gt = cat(t, q), train.py:352)."""

    def __init__(self, config: PWCLONetTrainerConfig = None, prediction_module=None, loss_module=None):
        self.config = config or PWCLONetTrainerConfig()
        c = self.config
        dev = torch.device(c.device)
        if dev.type != "cuda":
            raise RuntimeError("PWCLONetTrainer needs a CUDA device (there is no CPU path)")
        self.device = dev
        self.prediction_module_ = prediction_module or _PWCLONetPredictionModule(
            {"device": c.device, "num_points": c.num_points, "scalar_last": c.scalar_last})
        self.loss_module_ = loss_module or _PWCLONetLossModule(c.loss)
        self.prediction_module_.to(dev)
        self.loss_module_.to(dev)
        if c.optimizer_type != "adam":
            raise NotImplementedError("only the optimiser the PWCLO-Net recipe uses (adam) is built")
        self.arena = FlatArena([self.prediction_module_, self.loss_module_])
        self._optimizer = FlatAdam(self.arena, lr=c.optimizer_learning_rate, betas=(c.optimizer_beta, c.optimizer_momentum),
                                   weight_decay=c.optimizer_weight_decay)
        self._bn_scheduler = BNMomentumScheduler(
            self.prediction_module_, bn_lambda=lambda e: bn_momentum(e, c.bn_momentum_init, c.bn_decay_rate, c.bn_decay_step,
                                                                    c.bn_momentum_max), last_epoch=-1)
        self.num_epochs = 0
        self.train_iter = 0
        self.eval_iter = 0
        self.best = None
        self._average_train_loss = None
        self._graph = None
        self.broadcast_parameters()

    # ---- data-parallel plumbing
    def broadcast_parameters(self):
        """rank 0's initial weights and BN buffers to every rank (one broadcast of the arena + buffers)"""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.broadcast(self.arena.param, src=0)
            for b in self.prediction_module_.buffers():
                dist.broadcast(b, src=0)

    # ---- one iteration
    def pred_loss_forward_pass(self, batch, geoms=None):
        pred, pred_dict = self.prediction_module_(batch, geoms=geoms)
        gt = torch.cat((batch[3], batch[2]), 1)
        loss, log = self.loss_module_(pred, gt)
        return loss, log, pred

    def _step_body(self, batch, geoms=None):
        collect = os.environ.get("PWCLO_GRAD_COLLECT", "1") != "0"
        if collect:
            self.arena.detach_grads()
        else:
            self._optimizer.zero_grad()
        loss, log, pred = self.pred_loss_forward_pass(batch, geoms)
        loss.backward()
        if collect:
            self.arena.collect_grads()
        scale = all_reduce_gradients(self.arena)
        self._optimizer.step(grad_scale=scale)
        return loss.detach(), log, pred.detach()

    def train_step(self, batch):
        self.prediction_module_.train()
        self.loss_module_.train()
        out = self._step_body(batch)
        self.train_iter += 1
        return out

    # ---- the whole step as ONE CUDA graph (the eager step is ~2 500 launches and bound by the host)
    def capture(self, batch, warmup=3, prefetch=None):
        """Capture zero_grad -> forward -> loss -> backward -> all-reduce -> Adam for batches shaped like
        `batch` (its first four entries are copied into static buffers).  The `warmup` eager steps it runs first
        ARE training steps.  The graph stays valid across steps and learning-rate changes (step counter and lr
        live in device memory); it is dropped when the BN momentum changes (a captured scalar) or on train(False).

        prefetch (default on, PWCLO_GEO_PREFETCH=0 disables): the coordinates-only part of the forward -- the sampling
        chain and neighbour searches of the pyramid, ~3 ms of a 17 ms step at 8 pairs x 16 384 points, and independent of
        the weights -- is captured as its OWN graph; train_step_graphed(batch, next_batch=...) replays it for the next
        batch on a second stream while the current step is still running."""
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        # Data parallel: the NCCL all-reduce of the gradient arena is captured INSIDE the step graph (NCCL supports stream
        # capture); every rank must capture and replay in lock step.  Measured on 2 B200: 18.0 ms per step against
        # 55 ms for the eager data-parallel step, the captured all-reduce 31 us.  What hung in round 1 was process
        # teardown with the graph still alive: call drop_graph() (then synchronize + barrier) before
        # destroy_process_group(), see close().  capture_error_mode="thread_local": NCCL's watchdog thread may
        # touch the CUDA API while this thread captures.
        if prefetch is None:
            prefetch = os.environ.get("PWCLO_GEO_PREFETCH", "1") != "0"
        self.prediction_module_.train()
        self.loss_module_.train()
        self._static_batch = [b.to(self.device).clone() if torch.is_tensor(b) else b for b in batch[:4]]
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._step_body(self._static_batch)
                self.train_iter += 1
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        kw = {"capture_error_mode": "thread_local"} if multi else {}
        self._geo_graph, self._geo_cur, self._prefetched, self._copied = None, None, None, None
        if prefetch:
            # graph 1: clouds of the NEXT batch (static copies) -> its pyramid geometry (static outputs of the graph)
            self._geo_stream = torch.cuda.Stream(device=self.device)
            self._geo_done = torch.cuda.Event()
            self._geo_in = [self._static_batch[0].clone(), self._static_batch[1].clone()]
            self._geo_graph = torch.cuda.CUDAGraph()
            with torch.no_grad():
                with torch.cuda.graph(self._geo_graph, **kw):
                    self._geo_out = self.prediction_module_.geometry(self._geo_in)
                self._geo_graph.replay()
                torch.cuda.synchronize(self.device)
                # the step graph reads its geometry from these buffers; a replay of graph 1 must not overwrite them under it
                self._geo_cur = [tuple(t.clone() for t in g) for g in self._geo_out]
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph, **kw):
            self._static_out = self._step_body(self._static_batch, self._geo_cur)
        self.train_iter += 1                          # capture executes nothing, but the step below replays it once
        self._graph.replay()
        self._graph_momentum = self._bn_scheduler.last_momentum
        return self._static_out

    def _geometry_into_cur(self, batch):
        """make self._geo_cur hold the pyramid geometry of `batch`: from the prefetch if it was issued for this very
        batch object, else by replaying the geometry graph now on the current stream"""
        cur = torch.cuda.current_stream(self.device)
        if self._prefetched is not None and self._prefetched is batch:
            cur.wait_event(self._geo_done)
        else:
            if self._prefetched is not None:
                cur.wait_event(self._geo_done)          # an unused prefetch still owns the buffers until it has finished
            self._geo_in[0].copy_(batch[0], non_blocking=True)
            self._geo_in[1].copy_(batch[1], non_blocking=True)
            self._geo_graph.replay()
        self._prefetched = None
        torch._foreach_copy_([t for g in self._geo_cur for t in g], [t for g in self._geo_out for t in g])
        self._copied = torch.cuda.Event()
        self._copied.record(cur)

    def train_step_graphed(self, batch, next_batch=None):
        """copy the batch into the static buffers and replay the captured step; returns the (static) loss, log
        and prediction tensors, valid until the next replay.  next_batch: the batch the NEXT call will receive (the same
        object): its coordinates-only work is replayed on a second stream while this step runs."""
        if getattr(self, "_graph", None) is None:
            self.capture(batch)
            # fall through: the capture replayed one step on `batch`; continue with the prefetch logic only
            if next_batch is not None and self._geo_graph is not None:
                self._issue_prefetch(next_batch)
            return self._static_out
        if self._geo_graph is not None:
            self._geometry_into_cur(batch)
        for dst, src in zip(self._static_batch, batch):
            if torch.is_tensor(dst):
                dst.copy_(src, non_blocking=True)
        self._graph.replay()
        self.train_iter += 1
        if next_batch is not None and self._geo_graph is not None:
            self._issue_prefetch(next_batch)
        return self._static_out

    def _issue_prefetch(self, next_batch):
        """replay the geometry graph for `next_batch` on the geometry stream.  It overwrites the graph's static outputs, so
        it must come after this step's copy of them into the step graph's own buffers (`_copied`, recorded right after that
        copy and BEFORE the step graph on the current stream) -- which lets it run beside the whole step."""
        ev = getattr(self, "_copied", None)
        if ev is None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
        gs = self._geo_stream
        gs.wait_event(ev)
        with torch.cuda.stream(gs), torch.no_grad():
            self._geo_in[0].copy_(next_batch[0], non_blocking=True)
            self._geo_in[1].copy_(next_batch[1], non_blocking=True)
            self._geo_graph.replay()
            self._geo_done.record(gs)
        self._prefetched = next_batch

    def drop_graph(self):
        self._graph = None
        self._static_out = None
        self._geo_graph = None
        self._geo_out = None
        self._geo_cur = None
        self._prefetched = None

    def close(self):
        """release the captured step before the process group goes away: a live CUDA graph that holds NCCL work makes
        destroy_process_group() / interpreter exit hang (measured, round 1).  Collective: every rank calls it."""
        self.drop_graph()
        self._static_batch = None
        torch.cuda.synchronize(self.device)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.barrier()
            torch.cuda.synchronize(self.device)

    def train_epoch(self, batches):
        """`batches`: iterable of collated batches already on the device (or host tensors, which are
        sent with non_blocking copies).  Returns the mean loss of the epoch (one D2H read at the end)."""
        total, n = torch.zeros((), device=self.device), 0
        for batch in batches:
            batch = [b.to(self.device, non_blocking=True) if torch.is_tensor(b) else b for b in batch]
            loss, _, _ = self.train_step(batch)
            total += loss
            n += 1
        self.end_epoch()
        mean = float(total) / max(n, 1)
        if not math.isfinite(mean):
            raise RuntimeError("\n[ERROR] Loss is NaN.\n")
        self._average_train_loss = mean
        return mean

    def end_epoch(self):
        """scheduler stepping of ATrainer.train (trainer.py:469-483)"""
        c = self.config
        e = self.num_epochs + 1
        if c.coslr:
            self._optimizer.lr = cosine_lr(c.optimizer_learning_rate, e, c.num_epochs, c.scheduler_decay_clip)
        else:
            self._optimizer.lr = exponential_lr(c.optimizer_learning_rate, e, c.optimizer_scheduler_decay, c.scheduler_decay_clip)
        self._bn_scheduler.step(self.num_epochs)
        if getattr(self, "_graph", None) is not None and self._bn_scheduler.last_momentum != self._graph_momentum:
            self.drop_graph()                         # BN momentum is a captured scalar: re-capture on the next step
        self.num_epochs = e

    # ---- checkpoints (trainer.py:840-907)
    def save_checkpoint(self, checkpoint_file):
        if not checkpoint_file:
            return
        sd = {"optimizer": self._optimizer.state_dict(),
              "loss_module": self.loss_module_.state_dict(),
              "prediction_module": self.prediction_module_.state_dict(),
              "num_train_epochs": self.num_epochs, "train_iter": self.train_iter, "eval_iter": self.eval_iter,
              "best": self.best, "last_lr": [self._optimizer.lr]}
        if self._average_train_loss is not None:
            sd["average_train_loss"] = self._average_train_loss
        torch.save(sd, checkpoint_file)

    def load_checkpoint(self, checkpoint_file, fail_if_absent=True):
        import os
        if not os.path.exists(checkpoint_file):
            if fail_if_absent:
                raise FileNotFoundError(checkpoint_file)
            return False
        sd = torch.load(checkpoint_file, map_location=self.device, weights_only=False)
        # load_state_dict copies INTO the arena views, so the flat layout survives
        self.prediction_module_.load_state_dict(sd["prediction_module"])
        self.loss_module_.load_state_dict(sd["loss_module"])
        if "optimizer" in sd:
            self._optimizer.load_state_dict(sd["optimizer"])
        self.num_epochs = sd.get("num_train_epochs", 0)
        self.train_iter = sd.get("train_iter", 0)
        self.eval_iter = sd.get("eval_iter", 0)
        self.best = sd.get("best")
        self._average_train_loss = sd.get("average_train_loss")
        if "last_lr" in sd:
            self._optimizer.lr = float(sd["last_lr"][0] if isinstance(sd["last_lr"], (list, tuple)) else sd["last_lr"])
        self._bn_scheduler.step(max(self.num_epochs - 1, 0))
        return True
