"""CPU suite for the training-side rows: the loss oracle against the
unmodified reference and its golden vectors, the Adam oracle against torch.optim.Adam, the flat
gradient arena with its one all-reduce (world_size 2, gloo), the schedules."""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from oracle import ref_shim, train_port
from pwclonet_pylidarslam_b200 import training as T
from tests import _common as C


def _golden_cases():
    g = dict(np.load(os.path.join(C.GOLD_DIR, "loss_kat.npz")))
    for i in range(int(g["n_cases"])):
        yield i, {k[len(f"c{i}_"):]: v for k, v in g.items() if k.startswith(f"c{i}_")}


def test_loss_oracle_matches_reference_golden():
    """bit-exact: the oracle is the same fp32 torch expression tree as loss_modules.py:424-544"""
    for i, c in _golden_cases():
        o = train_port.pose_loss(c["pred"], c["gt"], c["s"], bool(c["with_exp"]))
        np.testing.assert_array_equal(o["terms"][:13], c["terms"], err_msg=f"case {i}")
        np.testing.assert_array_equal(o["grad_pred"], c["grad_pred"], err_msg=f"case {i}")
        if bool(c["with_exp"]):
            np.testing.assert_array_equal(o["grad_s"], c["grad_s"], err_msg=f"case {i}")


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted")
def test_loss_oracle_matches_unmodified_reference_live():
    rng = np.random.default_rng(3)
    pred = rng.standard_normal((9, 4, 7)).astype(np.float32)
    gt = rng.standard_normal((9, 7)).astype(np.float32)
    mod = ref_shim.load_reference_loss(True, (0.2, -2.0))
    p = torch.tensor(pred, requires_grad=True)
    loss, log = mod(p, torch.tensor(gt))
    loss.backward()
    o = train_port.pose_loss(pred, gt, [0.2, -2.0])
    assert float(loss) == o["loss"]
    np.testing.assert_array_equal(p.grad.numpy(), o["grad_pred"])
    np.testing.assert_array_equal(mod.exp_weighting.s_param.grad.numpy(), o["grad_s"])
    # the product module exposes the same log keys as the reference
    ours = T._PWCLONetLossModule(T.PWCLONetLossConfig())
    assert set(log.keys()) == {f"loss_rot_l{i}" for i in range(1, 5)} | {f"loss_trans_l{i}" for i in range(1, 5)} | \
        {f"s_rot_l{i}" for i in range(1, 5)} | {f"s_trans_l{i}" for i in range(1, 5)} | \
        {f"loss_l{i}" for i in range(1, 5)} | {"loss", "s_param_trans", "s_param_rot"}
    assert list(ours.state_dict().keys()) == list(mod.state_dict().keys()) == ["exp_weighting.s_param"]


def test_loss_module_refuses_cpu():
    m = T._PWCLONetLossModule(T.PWCLONetLossConfig())
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(2, 4, 7), torch.zeros(2, 7))


def test_adam_oracle_matches_torch_adam():
    rng = np.random.default_rng(0)
    p0 = rng.standard_normal(1000).astype(np.float32)
    tp = nn.Parameter(torch.tensor(p0))
    opt = torch.optim.Adam([tp], lr=1e-3, betas=(0.9, 0.999), weight_decay=1e-3, foreach=False)
    p, m, v = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    for step in range(1, 8):
        g = rng.standard_normal(1000).astype(np.float32)
        tp.grad = torch.tensor(g)
        opt.step()
        p, m, v = train_port.adam_step(p, g, m, v, step, 1e-3, wd=1e-3)
        np.testing.assert_allclose(p, tp.detach().numpy(), rtol=2e-6, atol=1e-7)
    st = opt.state[tp]
    np.testing.assert_allclose(m, st["exp_avg"].numpy(), rtol=2e-6, atol=1e-7)   # cancellation in g - m
    np.testing.assert_allclose(v, st["exp_avg_sq"].numpy(), rtol=2e-6, atol=1e-12)


def test_schedules_match_torch_and_reference_formulas():
    p = nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([p], lr=1e-3)
    sch = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=100, eta_min=1e-6)
    for e in range(1, 40):
        opt.step()
        sch.step()
        assert math.isclose(sch.get_last_lr()[0], T.cosine_lr(1e-3, e, 100, 1e-6), rel_tol=1e-9)
    assert T.exponential_lr(1e-3, 3, 0.7, 1e-5) == 1e-3 * 0.7 ** 3
    assert T.exponential_lr(1e-3, 50, 0.7, 1e-5) == 1e-5
    # train.py:318-322 with the defaults of train.py:213-216
    assert [T.bn_momentum(e) for e in (0, 3, 4, 8, 40)] == [0.5, 0.5, 0.75, 0.875, 0.99]


def _tiny():
    torch.manual_seed(0)
    return nn.Sequential(nn.Linear(5, 7), nn.ReLU(), nn.Linear(7, 3))


def test_flat_arena_views_and_gradients():
    net, extra = _tiny(), T.ExponentialWeights(2, [0.0, -2.5])
    ref = [p.detach().clone() for p in list(net.parameters()) + list(extra.parameters())]
    arena = T.FlatArena([net, extra])
    assert arena.count == sum(r.numel() for r in ref) and arena.numel % 4 == 0
    for p, r, o in zip(arena.params, ref, arena.offsets):
        assert o % 4 == 0 and torch.equal(p, r)
        assert p.data_ptr() == arena.param.data_ptr() + 4 * o and p.grad.data_ptr() == arena.grad.data_ptr() + 4 * o
    x = torch.randn(4, 5)
    (net(x).sum() * extra.s_param.sum()).backward()
    g1 = arena.grad.clone()
    assert float(g1.abs().sum()) > 0                      # autograd accumulated INTO the arena
    for p, o in zip(arena.params, arena.offsets):
        assert p.grad.data_ptr() == arena.grad.data_ptr() + 4 * o
    arena.zero_grad()
    assert float(arena.grad.abs().sum()) == 0
    # load_state_dict writes through the views
    net.load_state_dict({k: torch.ones_like(v) for k, v in net.state_dict().items()})
    assert float(arena.param[:35].min()) == 1.0


def _dp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        net = _tiny()
        arena = T.FlatArena([net])
        torch.manual_seed(100 + rank)                     # each rank its own shard of the batch
        x = torch.randn(6, 5)
        arena.zero_grad()
        net(x).square().mean().backward()
        scale = T.all_reduce_gradients(arena)
        np.save(out % rank, np.concatenate([[scale], (arena.grad * scale).numpy(), x.numpy().reshape(-1)]))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_full_batch(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "r%d.npy")
    mp.spawn(_dp_worker, args=(2, port, out), nprocs=2, join=True)
    r0, r1 = np.load(out % 0), np.load(out % 1)
    assert r0[0] == 0.5 and r1[0] == 0.5
    net = _tiny()
    arena = T.FlatArena([net])
    n = arena.numel
    np.testing.assert_array_equal(r0[1:1 + n], r1[1:1 + n])          # both ranks hold the same averaged gradient
    x = torch.tensor(np.concatenate([r0[1 + n:].reshape(6, 5), r1[1 + n:].reshape(6, 5)]), dtype=torch.float32)
    net(x).square().mean().backward()                                  # the full batch on one process
    np.testing.assert_allclose(r0[1:1 + n], arena.grad.numpy(), rtol=1e-5, atol=1e-7)


def test_no_cpu_path_anywhere_in_the_new_rows():
    """trainer, odometry adapter, streaming pipeline, optimiser and pose post-processing refuse CPU devices / tensors
    instead of falling back"""
    from pwclonet_pylidarslam_b200 import odometry as O
    from pwclonet_pylidarslam_b200 import sharding
    with pytest.raises(RuntimeError, match="no CPU path"):
        T.PWCLONetTrainer(T.PWCLONetTrainerConfig(device="cpu"))
    with pytest.raises(RuntimeError, match="no CPU path"):
        O.PWCLONetOdometry({"device": "cpu"})
    with pytest.raises(RuntimeError, match="no CPU path"):
        sharding.PosePipeline(nn.Linear(2, 2), 1, 8)
    with pytest.raises(RuntimeError, match="no CPU path"):
        T.FlatAdam(T.FlatArena([_tiny()]))
    with pytest.raises(RuntimeError, match="no CPU path"):
        O.pose_params_to_matrices(torch.zeros(2, 4, 7))
    with pytest.raises(RuntimeError, match="no CPU path"):
        O.convert_to_absolute(torch.zeros(2, 4, 4, dtype=torch.float64))


def test_channel_padded_conv_is_the_same_function():
    """pytorch_utils.conv1x1_aligned (zero-padded input channels so the library picks aligned GEMM kernels in training) is
    conv(x) with the same input / weight gradients, for Conv2d and Conv1d and for every odd channel count of the model"""
    import torch.nn as nn
    from pwclonet_pylidarslam_b200 import pytorch_utils as pt
    old = pt.PAD_CONV_MIN_POSITIONS
    pt.PAD_CONV_MIN_POSITIONS = 1
    try:
        torch.manual_seed(3)
        for ci, mod, shape in ((19, nn.Conv2d, (2, 19, 8, 4)), (10, nn.Conv2d, (2, 10, 8, 4)), (67, nn.Conv1d, (2, 67, 16)),
                               (138, nn.Conv2d, (1, 138, 4, 2)), (64, nn.Conv2d, (2, 64, 4, 4))):
            conv = mod(ci, 8, 1, bias=False)
            x = torch.randn(shape, requires_grad=True)
            y0 = conv(x)
            g0 = torch.autograd.grad(y0.square().sum(), (x, conv.weight))
            y1 = pt.conv1x1_aligned(conv, x)
            g1 = torch.autograd.grad(y1.square().sum(), (x, conv.weight))
            assert torch.allclose(y0, y1, rtol=1e-6, atol=1e-6)
            for a, b in zip(g0, g1):
                assert a.shape == b.shape and torch.allclose(a, b, rtol=1e-5, atol=1e-5)
    finally:
        pt.PAD_CONV_MIN_POSITIONS = old


def test_concatenation_writes_the_zero_channels_for_its_consumer(monkeypatch):
    """pytorch_utils.cat_for: when the consumer's first convolution runs on zero-padded channels, the concatenation
    carries them (no padding copy), conv1x1_aligned takes the input as it is, and the result is conv(cat(...));
    off the fused training path it is a plain torch.cat"""
    import torch.nn as nn
    from pwclonet_pylidarslam_b200 import pytorch_utils as pt
    torch.manual_seed(5)
    mlp = pt.SharedMLP([19, 16, 16], bn=True).train()
    a, b = torch.randn(2, 3, 8, 4), torch.randn(2, 16, 8, 4, requires_grad=True)
    assert torch.equal(pt.cat_for(mlp, (a, b)), torch.cat((a, b), dim=1))
    monkeypatch.setattr(pt._Conv, "fused_train_path", lambda self, x: True)
    monkeypatch.setattr(pt, "PAD_CONV_MIN_POSITIONS", 1)
    x = pt.cat_for(mlp, (a, b))
    assert x.shape[1] == 20 and torch.equal(x[:, :19], torch.cat((a, b), dim=1)) and not x[:, 19:].any()
    conv = next(iter(mlp.children())).conv
    y = pt.conv1x1_aligned(conv, x)
    want = conv(torch.cat((a, b), dim=1))
    assert torch.allclose(y, want, rtol=1e-6, atol=1e-6)
    g = torch.autograd.grad(y.square().sum(), (b, conv.weight))
    gw = torch.autograd.grad(want.square().sum(), (b, conv.weight))
    for u, v in zip(g, gw):
        assert u.shape == v.shape and torch.allclose(u, v, rtol=1e-5, atol=1e-5)
    with pytest.raises(RuntimeError):
        pt.conv1x1_aligned(conv, torch.randn(2, 22, 8, 4))
    post = pt.SharedMLP([128, 64], bn=True).train()                  # aligned already: nothing appended
    assert pt.cat_for(post, (torch.randn(2, 64, 8, 1), torch.randn(2, 64, 8, 1))).shape[1] == 128
