"""Host-side mirror of the reference's P2/pytorch_utils.py for the PWCLO-Net path: `knn_point`
(now a single sm_100a kernel instead of materialised [B,S,N,3] tensors + topk) and the small
nn.Module wrappers whose *parameter names* define the checkpoint layout:
  SharedMLP.layer{i} -> Conv2d.{conv, bn.bn, activation}      (P2/pytorch_utils.py:52-83,114-167,236-269)
  Conv1d.{conv}                                               (P2/pytorch_utils.py:170-203)
"""
from typing import List

import torch
import torch.nn as nn

from . import _ext


def knn_point(nsample, xyz, new_xyz):
    """P2/pytorch_utils.py:32-49.  xyz [B,N,3] (reference set), new_xyz [B,S,3] (queries) ->
    (idx, idx) int32 [B,S,nsample]; the reference returns the index tensor in both slots (its
    "dist" output is overwritten by the indices, :46-47) and so do we."""
    idx = _ext.knn(xyz.contiguous(), new_xyz.contiguous(), nsample)
    return idx, idx


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
