"""Host-side mirror of the reference's P2/pytorch_utils.py for the PWCLO-Net path: `knn_point`
(now a single sm_100a kernel instead of materialised [B,S,N,3] tensors + topk) and the small
nn.Module wrappers whose *parameter names* define the checkpoint layout:
  SharedMLP.layer{i} -> Conv2d.{conv, bn.bn, activation}      (P2/pytorch_utils.py:52-83,114-167,236-269)
  Conv1d.{conv}                                               (P2/pytorch_utils.py:170-203)
"""
from typing import List

import os

import torch
import torch.nn as nn

from . import _ext


def knn_point(nsample, xyz, new_xyz):
    """P2/pytorch_utils.py:32-49.  xyz [B,N,3] (reference set), new_xyz [B,S,3] (queries) ->
    (idx, idx) int32 [B,S,nsample]; the reference returns the index tensor in both slots (its
    "dist" output is overwritten by the indices, :46-47) and so do we."""
    idx = _ext.knn(xyz.contiguous(), new_xyz.contiguous(), nsample)
    return idx, idx


class FusedBNReLUTrain(torch.autograd.Function):
    """Train-mode BatchNorm (batch statistics, running-statistics update) + ReLU of one shared-MLP layer on the
    sm_100a kernels (`pwclo_bn_relu_train_fwd/_bwd`): what nn.BatchNorm2d / nn.BatchNorm1d followed by nn.ReLU
    compute (P2/pytorch_utils.py:86-167), in two launches per direction instead of cuDNN's batch norm plus separate
    ReLU kernels."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, momentum, eps):
        import ctypes
        from . import _lib
        x = x.contiguous()
        B, C = x.shape[0], x.shape[1]
        HW = x.numel() // (B * C)
        L = _lib.lib()
        p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None       # running_* may be None: no update
        y = torch.empty_like(x)
        mean = torch.empty(C, dtype=torch.float32, device=x.device)
        invstd = torch.empty(C, dtype=torch.float32, device=x.device)
        ws = torch.empty(L.pwclo_bn_relu_workspace_bytes(B, C, HW) // 8, dtype=torch.float64, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(L.pwclo_bn_relu_train_fwd(p(x), p(weight), p(bias), B, C, HW, float(eps), float(momentum), p(running_mean),
                                                 p(running_var), p(y), p(mean), p(invstd), p(ws), _lib.stream_ptr()),
                       "bn_relu_train_fwd")
        ctx.save_for_backward(x, weight, bias, mean, invstd)
        ctx.dims = (B, C, HW)
        FusedBNReLUTrain.last_stats = (mean, invstd, B * HW)
        return y

    @staticmethod
    def backward(ctx, dy):
        import ctypes
        from . import _lib
        x, weight, bias, mean, invstd = ctx.saved_tensors
        B, C, HW = ctx.dims
        dy = dy.contiguous()
        L = _lib.lib()
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        dx = torch.empty_like(x)
        dgamma = torch.empty_like(weight)
        dbeta = torch.empty_like(bias)
        ws = torch.empty(L.pwclo_bn_relu_workspace_bytes(B, C, HW) // 8, dtype=torch.float64, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(L.pwclo_bn_relu_train_bwd(p(x), p(dy), p(weight), p(bias), p(mean), p(invstd), B, C, HW, p(dx), p(dgamma),
                                                 p(dbeta), p(ws), _lib.stream_ptr()), "bn_relu_train_bwd")
        return dx, dgamma, dbeta, None, None, None, None


class SkinnyConv1x1(torch.autograd.Function):
    """1x1 convolution without bias for the few-channel layers of the first set conv (6 -> 8 -> 8 -> 16 on ~5e5
    positions) in training: streaming forward / input gradient (`pwclo_conv1x1_small`) and a deterministic two-launch
    weight gradient (`pwclo_conv1x1_wgrad`) instead of library GEMMs whose tiles are 90 % padding (the weight gradient,
    an 8 x 6 GEMM with K = 524 288, took 0.67 ms per call)."""
    FWD = {(6, 8), (8, 8), (8, 16), (3, 8)}
    DX = {(8, 8), (8, 16)}          # forward (CI, CO) whose input gradient kernel (CO -> CI) exists

    @staticmethod
    def usable(x, conv, shape=None):
        """`shape`: the input's shape when `x` is only a tensor of the same device / dtype"""
        w = conv.weight
        shape = tuple(x.shape if shape is None else shape)
        n = 1
        for d in shape:
            n *= d
        return (x.is_cuda and x.dtype == torch.float32 and conv.bias is None and w.shape[2:].numel() == 1
                and (w.shape[1], w.shape[0]) in SkinnyConv1x1.FWD and (n // (shape[0] * shape[1])) % 4 == 0
                and n // shape[1] >= 65536)

    @staticmethod
    def forward(ctx, x, w):
        import ctypes
        from . import _lib
        x = x.contiguous()
        B, CI = x.shape[0], x.shape[1]
        CO = w.shape[0]
        HW = x.numel() // (B * CI)
        y = torch.empty((B, CO) + tuple(x.shape[2:]), dtype=x.dtype, device=x.device)
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        wc = w.contiguous()
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().pwclo_conv1x1_small(p(x), p(wc), 0, B, CI, CO, HW, p(y), _lib.stream_ptr()), "conv1x1_small")
        ctx.save_for_backward(x, wc)
        return y

    @staticmethod
    def backward(ctx, dy):
        import ctypes
        from . import _lib
        x, w = ctx.saved_tensors
        B, CI = x.shape[0], x.shape[1]
        CO = w.shape[0]
        HW = x.numel() // (B * CI)
        dy = dy.contiguous()
        L = _lib.lib()
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        dw = torch.empty_like(w)
        ws = torch.empty(max(1, L.pwclo_conv1x1_wgrad_workspace_bytes(B, CI, CO, HW) // 4), dtype=torch.float32, device=x.device)
        dx = None
        with torch.cuda.device(x.device):
            _lib.check(L.pwclo_conv1x1_wgrad(p(x), p(dy), B, CI, CO, HW, p(dw), p(ws), _lib.stream_ptr()), "conv1x1_wgrad")
            if ctx.needs_input_grad[0]:
                if (CI, CO) in SkinnyConv1x1.DX:
                    dx = torch.empty_like(x)
                    _lib.check(L.pwclo_conv1x1_small(p(dy), p(w), 1, B, CO, CI, HW, p(dx), _lib.stream_ptr()), "conv1x1_small")
                else:
                    dx = torch.einsum("oc,bo...->bc...", w.reshape(CO, CI), dy)
        return dx, dw


class MaxPoolLastDim(torch.autograd.Function):
    """max over the last axis of [B, C, S, K] -> [B, C, S] on the sm_100a kernels (`pwclo_maxpool_lastdim_fwd/_bwd`): what
    F.max_pool2d(x, kernel_size=[1, K]).squeeze(-1) computes in the set conv / set upconv (P2/pointnet2_modules.py:239-243,
    :499-506), same tie rule (first maximum), with a backward that writes every gradient element once (ATen's
    max_pool_backward_nchw took 40 us per call on these shapes)."""

    @staticmethod
    def forward(ctx, x):
        import ctypes
        from . import _lib
        x = x.contiguous()
        K = x.shape[-1]
        rows = x.numel() // K
        y = torch.empty(x.shape[:-1], dtype=x.dtype, device=x.device)
        arg = torch.empty(x.shape[:-1], dtype=torch.uint8, device=x.device)
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().pwclo_maxpool_lastdim_fwd(p(x), rows, K, p(y), p(arg), _lib.stream_ptr()), "maxpool_lastdim_fwd")
        ctx.save_for_backward(arg)
        ctx.K = K
        ctx.mark_non_differentiable(arg)
        return y

    @staticmethod
    def backward(ctx, dy):
        import ctypes
        from . import _lib
        (arg,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty(tuple(dy.shape) + (ctx.K,), dtype=dy.dtype, device=dy.device)
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        with torch.cuda.device(dy.device):
            _lib.check(_lib.lib().pwclo_maxpool_lastdim_bwd(p(dy), p(arg), dy.numel(), ctx.K, p(dx), _lib.stream_ptr()),
                       "maxpool_lastdim_bwd")
        return dx


def max_over_neighbours(x):
    """[B, C, S, K] -> [B, C, S]: the reference's F.max_pool2d(x, kernel_size=[1, K]).squeeze(-1)"""
    if x.is_cuda and x.dtype == torch.float32 and x.shape[-1] <= 255 and os.environ.get("PWCLO_MAXPOOL", "1") != "0":
        return MaxPoolLastDim.apply(x)
    return torch.nn.functional.max_pool2d(x, kernel_size=[1, x.size(3)]).squeeze(-1)


class SoftmaxPool(torch.autograd.Function):
    """sum_k softmax(w, dim=-1)[..., k] * x[..., k] on the sm_100a kernels (`pwclo_softmax_pool_fwd/_bwd`): the attentive
    pooling of the cost volume (PW/costvolume.py:139-145, :181-188) as one pass forward and one backward; the softmax
    is recomputed from w in the backward, so only the two inputs are kept alive"""
    KS = (4, 6, 8, 16, 32)

    @staticmethod
    def forward(ctx, w, x):
        import ctypes
        from . import _lib
        w, x = w.contiguous(), x.contiguous()
        K = w.shape[-1]
        out = torch.empty(w.shape[:-1], dtype=w.dtype, device=w.device)
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        with torch.cuda.device(w.device):
            _lib.check(_lib.lib().pwclo_softmax_pool_fwd(p(w), p(x), w.numel() // K, K, p(out), _lib.stream_ptr()),
                       "softmax_pool_fwd")
        ctx.save_for_backward(w, x)
        return out

    @staticmethod
    def backward(ctx, g):
        import ctypes
        from . import _lib
        w, x = ctx.saved_tensors
        K = w.shape[-1]
        g = g.contiguous()
        dw, dx = torch.empty_like(w), torch.empty_like(x)
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        with torch.cuda.device(w.device):
            _lib.check(_lib.lib().pwclo_softmax_pool_bwd(p(w), p(x), p(g), w.numel() // K, K, p(dw), p(dx), _lib.stream_ptr()),
                       "softmax_pool_bwd")
        return dw, dx


def softmax_pool(w, x):
    """[B, C, S, K] x 2 -> [B, C, S]: torch.sum(F.softmax(w, dim=3) * x, dim=3) (the reference's expression, used as it is
    without autograd: the parity path of inference)"""
    if (w.is_cuda and w.dtype == torch.float32 and x.dtype == torch.float32 and w.shape == x.shape and torch.is_grad_enabled()
            and (w.requires_grad or x.requires_grad) and w.shape[-1] in SoftmaxPool.KS
            and os.environ.get("PWCLO_SOFTMAX_POOL", "1") != "0"):
        return SoftmaxPool.apply(w, x)
    return torch.sum(nn.functional.softmax(w, dim=3) * x, dim=3)


class deferred_bn_counters:
    """Context: the `num_batches_tracked += 1` of every train-mode BatchNorm executed inside it (93 one-element launches
    per training step of PWCLO-Net) is applied on exit with one multi-tensor add per distinct increment.  Same values as
    nn.BatchNorm's own bookkeeping (a layer that ran twice -- the two frames of a pair -- counts twice)."""
    pending = None

    def __enter__(self):
        self.outer = deferred_bn_counters.pending
        deferred_bn_counters.pending = {} if self.outer is None else self.outer
        return self

    def __exit__(self, *exc):
        if self.outer is None:
            todo, deferred_bn_counters.pending = deferred_bn_counters.pending, None
            by_count = {}
            for t, c in todo.values():
                by_count.setdefault(c, []).append(t)
            for c, ts in by_count.items():
                torch._foreach_add_(ts, c)
        return False

    @staticmethod
    def bump(counter):
        pend = deferred_bn_counters.pending
        if pend is None:
            counter.add_(1)
        else:
            rec = pend.setdefault(id(counter), [counter, 0])
            rec[1] += 1


class deferred_running_stats:
    """Context for a branch that runs CONCURRENTLY with another use of the same modules (the second frame's pyramid shares
    its SharedMLPs with the first frame's): the train-mode BatchNorm kernels inside it leave the running statistics alone
    and record their batch statistics instead; apply() performs  running = (1 - momentum) * running + momentum * batch
    (unbiased variance) for all of them with a few multi-tensor launches -- after the join, i.e. after the other branch's
    in-kernel updates, which is the reference's order (frame 1, then frame 2)."""
    active = None

    def __enter__(self):
        self.records = []
        self.outer, deferred_running_stats.active = deferred_running_stats.active, self.records
        return self

    def __exit__(self, *exc):
        deferred_running_stats.active = self.outer
        return False

    def apply(self):
        if not self.records:
            return
        with torch.no_grad():
            rm = [n.running_mean for n, *_ in self.records]
            rv = [n.running_var for n, *_ in self.records]
            mom = [float(n.momentum) for n, *_ in self.records]
            keep = [1.0 - m for m in mom]
            torch._foreach_mul_(rm, keep)
            torch._foreach_add_(rm, torch._foreach_mul([mean for _, mean, _, _ in self.records], mom))
            var = torch._foreach_reciprocal(torch._foreach_mul([i for _, _, i, _ in self.records], [i for _, _, i, _ in self.records]))
            torch._foreach_sub_(var, [float(n.eps) for n, *_ in self.records])                   # biased batch variance
            torch._foreach_mul_(var, [m * cnt / max(cnt - 1, 1) for m, (_, _, _, cnt) in zip(mom, self.records)])
            torch._foreach_mul_(rv, keep)
            torch._foreach_add_(rv, var)
        self.records.clear()


_BRANCH_STREAMS = {}


def branch_stream(device, which=0):
    """The second CUDA stream of the training composition (one per device): independent branches of the network -- the second
    frame's pyramid, the mask set-upconv, the second stage's geometry encoding of a cost volume -- run on it beside the main
    stream, forward and (autograd replays an op's backward on the stream of its forward) backward.  The layers are hundreds of
    small launches that do not fill the machine one at a time.  `which` selects one of several such streams (a cost volume
    forks while its caller's own branch may still be running).  Returns None when PWCLO_TRAIN_STREAMS=0 or on the CPU."""
    if device.type != "cuda" or os.environ.get("PWCLO_TRAIN_STREAMS", "1") == "0":
        return None
    key = (device.index if device.index is not None else torch.cuda.current_device(), which)
    if key not in _BRANCH_STREAMS:
        _BRANCH_STREAMS[key] = torch.cuda.Stream(device=device)
    return _BRANCH_STREAMS[key]


class on_branch:
    """with on_branch(stream): ...   forks `stream` from the current one; .join(*tensors) makes the current stream wait for
    the branch and tells the allocator that the tensors produced there are consumed here.  stream None: plain in-line code."""

    def __init__(self, stream):
        self.stream = stream

    def __enter__(self):
        if self.stream is not None:
            self.main = torch.cuda.current_stream(self.stream.device)
            self.stream.wait_stream(self.main)
            self.ctx = torch.cuda.stream(self.stream)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.stream is not None:
            self.ctx.__exit__(*exc)
        return False

    def join(self, *tensors):
        if self.stream is not None:
            self.main.wait_stream(self.stream)
            for t in tensors:
                if torch.is_tensor(t):
                    t.record_stream(self.main)


SKINNY_CONV = os.environ.get("PWCLO_SKINNY_CONV", "1") != "0"
FUSED_BN_RELU = os.environ.get("PWCLO_FUSED_BN", "1") != "0"
PAD_CONV = os.environ.get("PWCLO_PAD_CONV", "1") != "0"
PAD_CONV_MIN_POSITIONS = int(os.environ.get("PWCLO_PAD_CONV_MIN", "8192"))


def conv_pad_channels(conv, positions):
    """zero channels conv1x1_aligned appends to the input of `conv` (0: the convolution runs as it is)"""
    ci = conv.in_channels
    pad = (16 - ci) if ci < 16 else (-ci) % 4
    if (not PAD_CONV or pad == 0 or conv.bias is not None or conv.weight.shape[2:].numel() != 1
            or positions < PAD_CONV_MIN_POSITIONS or not isinstance(conv, (nn.Conv2d, nn.Conv1d))):
        return 0
    return pad


def conv1x1_aligned(conv, x):
    """conv(x) for a bias-free 1x1 convolution whose input-channel count is not a multiple of 4 (the concatenated inputs
    of PWCLO-Net: 19, 35, 67, 138, 74, 42, 10 channels), training on the GPU: input and weight are zero-padded along the
    channel axis (to 16 below 16 channels, else to the next multiple of 4).  The padded channels contribute exact zeros; what
    changes is the library kernel: with an odd leading dimension of W the weight / input gradient GEMMs fall back to
    `..._align1` and `gemmSN_TN` kernels (8 x 67 -> 128 on 131 072 positions: backward 331 us -> 80 us with 68 channels;
    19 -> 16 on 262 144 positions: 302 -> 87 us; profiles/round2_conv_pad.txt).  An input that already carries the zero
    channels (`cat_for`) is used as it is."""
    ci = conv.in_channels
    have = x.shape[1] - ci
    pad = conv_pad_channels(conv, x.numel() // x.shape[1])
    if have == 0 and pad == 0:
        return conv(x)
    if have not in (0, pad):
        raise RuntimeError(f"conv1x1_aligned: input with {x.shape[1]} channels for a convolution of {ci} (+{pad})")
    spec = (0, 0) * (x.dim() - 2) + (0, pad)
    xp = x if have else nn.functional.pad(x, spec)
    wp = nn.functional.pad(conv.weight, spec)
    return nn.functional.conv2d(xp, wp) if x.dim() == 4 else nn.functional.conv1d(xp, wp)


def cat_for(consumer, tensors):
    """torch.cat(tensors, dim=1) as the input of `consumer` (a Conv2d / Conv1d unit of this module or a SharedMLP): when
    the consumer's first convolution is going to run on zero-padded channels (conv1x1_aligned), the zero channels are
    written by this concatenation, which saves the padding copy of the largest tensors of the step and its backward"""
    first = consumer
    while isinstance(first, nn.Sequential) and not isinstance(first, _Conv):
        first = next(iter(first.children()))
    x0 = tensors[0]
    if isinstance(first, _Conv) and first.fused_train_path(x0):
        ci = sum(t.shape[1] for t in tensors)
        shape = (x0.shape[0], ci) + tuple(x0.shape[2:])
        if ci == first.conv.in_channels and not (SKINNY_CONV and SkinnyConv1x1.usable(x0, first.conv, shape)):
            pad = conv_pad_channels(first.conv, x0.numel() // x0.shape[1])
            if pad:
                tensors = tuple(tensors) + (x0.new_zeros((x0.shape[0], pad) + tuple(x0.shape[2:])),)
    return torch.cat(tuple(tensors), dim=1)


class _BN(nn.Sequential):
    """P2/pytorch_utils.py:86-105: wrapper holding the norm layer under the name `bn`."""

    def __init__(self, size, norm):
        super().__init__()
        self.add_module("bn", norm(size))
        nn.init.constant_(self[0].weight, 1.0)
        nn.init.constant_(self[0].bias, 0)


class _Conv(nn.Sequential):
    """conv -> (bn) -> (activation), post-activation order only (preact is never used by PWCLO-Net)."""

    def __init__(self, conv, norm, in_size, out_size, activation, bn, init, bias):
        super().__init__()
        bias = bias and (not bn)
        unit = conv(in_size, out_size, kernel_size=1, stride=1, padding=0, bias=bias)
        init(unit.weight)
        if bias:
            nn.init.constant_(unit.bias, 0)
        self.add_module("conv", unit)
        if bn:
            self.add_module("bn", _BN(out_size, norm))
        if activation is not None:
            self.add_module("activation", activation)

    def fused_train_path(self, x):
        """training on the GPU with BatchNorm + ReLU: conv by torch, then ONE fused BN(train) + ReLU op"""
        bn = getattr(self, "bn", None)
        if not (FUSED_BN_RELU and self.training and bn is not None and isinstance(getattr(self, "activation", None), nn.ReLU)
                and x.is_cuda and x.dtype == torch.float32):
            return False
        norm = bn[0]
        return bool(norm.track_running_stats and norm.momentum is not None and norm.affine)

    def forward(self, x):
        """training on the GPU with BatchNorm + ReLU: conv by torch, then ONE fused BN(train) + ReLU op (same
        parameters / buffers / state-dict keys: the nn.BatchNorm module stays the owner of its tensors)"""
        if self.fused_train_path(x):
            norm = self.bn[0]
            y = (SkinnyConv1x1.apply(x, self.conv.weight) if SKINNY_CONV and SkinnyConv1x1.usable(x, self.conv)
                 else conv1x1_aligned(self.conv, x))
            deferred_bn_counters.bump(norm.num_batches_tracked)
            later = deferred_running_stats.active
            if later is None:
                return FusedBNReLUTrain.apply(y, norm.weight, norm.bias, norm.running_mean, norm.running_var, norm.momentum,
                                              norm.eps)
            out = FusedBNReLUTrain.apply(y, norm.weight, norm.bias, None, None, norm.momentum, norm.eps)
            later.append((norm,) + FusedBNReLUTrain.last_stats)
            return out
        return super().forward(x)


class Conv2d(_Conv):
    def __init__(self, in_size, out_size, *, activation=nn.ReLU(inplace=True), bn=False,
                 init=nn.init.kaiming_normal_, bias=True, **_ignored):
        super().__init__(nn.Conv2d, nn.BatchNorm2d, in_size, out_size, activation, bn, init, bias)


class Conv1d(_Conv):
    def __init__(self, in_size, out_size, *, kernel_size=1, padding=0, activation=nn.ReLU(inplace=True), bn=False,
                 init=nn.init.kaiming_normal_, bias=True, **_ignored):
        assert kernel_size == 1
        super().__init__(nn.Conv1d, nn.BatchNorm1d, in_size, out_size, activation, bn, init, bias)


class SharedMLP(nn.Sequential):
    def __init__(self, args: List[int], *, bn=False, activation=nn.ReLU(inplace=True), init=nn.init.kaiming_normal_,
                 name=""):
        super().__init__()
        for i in range(len(args) - 1):
            self.add_module(name + f"layer{i}", Conv2d(args[i], args[i + 1], bn=bn, activation=activation, init=init))


def set_bn_momentum_default(bn_momentum):
    """P2/pytorch_utils.py:310-316: the function BNMomentumScheduler applies to every module"""
    def fn(m):
        if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d)):
            m.momentum = bn_momentum
    return fn


class BNMomentumScheduler(object):
    """P2/pytorch_utils.py:319-347."""

    def __init__(self, model, bn_lambda, last_epoch=-1):
        if not isinstance(model, nn.Module):
            raise RuntimeError(f"Class '{type(model).__name__}' is not a PyTorch nn Module")
        self.model, self.lmbd = model, bn_lambda
        self.last_momentum = bn_lambda(0)
        self.step(last_epoch + 1)
        self.last_epoch = last_epoch

    def step(self, epoch=None):
        if epoch is None:
            epoch = self.last_epoch + 1
        self.last_epoch = epoch
        self.last_momentum = self.lmbd(epoch)
        m = self.last_momentum

        def fn(mod):
            if isinstance(mod, (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d)):
                mod.momentum = m

        self.model.apply(fn)
