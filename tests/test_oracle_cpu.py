"""CPU suite (-m "not gpu"): the oracle against the reference's golden vectors, host logic, and the
C-ABI surface.  No GPU compute."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import cpu_ops, ref_shim
from oracle.pwclo_port import Port
from tests import _common as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ops_known_answers():
    g = dict(np.load(os.path.join(C.GOLD_DIR, "ops_kat.npz")))
    x = g["xyz"]
    q = x[:, :77] + 0.01
    np.testing.assert_array_equal(cpu_ops.fps(x, 150), g["fps_150"])
    np.testing.assert_array_equal(cpu_ops.fps(x, 150, False, 1024), g["fps_150_cap1024_noskip"])
    np.testing.assert_array_equal(cpu_ops.knn(x, q, 8, 0), g["knn8_order0"])
    np.testing.assert_array_equal(cpu_ops.knn(x, q, 8, 1), g["knn8_order1"])
    np.testing.assert_array_equal(cpu_ops.ball_query(q, x, 1.0, 8), g["ball_r1_8"])
    d2, i3 = cpu_ops.three_nn(q, x)
    np.testing.assert_array_equal(i3, g["three_nn_idx"])
    np.testing.assert_array_equal(d2, g["three_nn_d2"])


def test_fps_properties():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 500, 3)).astype(np.float32)
    i = cpu_ops.fps(x, 500)
    assert (i[:, 0] == 0).all()
    for b in range(2):
        assert sorted(i[b].tolist()) == list(range(500))      # a permutation when m == n and no ties
    # the origin ball is never selected (except the forced start index 0)
    x[0, 10] = 0.0
    assert 10 not in cpu_ops.fps(x, 499)[0].tolist()
    # tie order: two identical far points, the tree prefers the smaller bit-reversed slot
    y = np.zeros((1, 8, 3), np.float32) + 1.0
    y[0, 3] = y[0, 6] = [5, 5, 5]
    # T = 8: bitrev3(3)=6, bitrev3(6)=3 -> index 6 wins the tie
    assert cpu_ops.fps(y, 2)[0, 1] == 6


def test_fps_prefix_property_holds_in_the_checker():
    """SURVEY 0.8, the fact pwclo_furthest_point_sampling_prefix rests on: FPS of the FPS-ordered output of a tie-free
    run returns 0..m-1 (restated reference kernel, EXT/src/sampling_gpu.cu:85-172), at every nested level of the pyramid;
    duplicated points (exact ties) break it, which is why the kernels carry a tie flag."""
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((2, 4096, 3)) * np.array([20, 2, 20])).astype(np.float32)
    cur = x
    for m in (1024, 512, 128, 32):
        i = cpu_ops.fps(cur, m)
        if cur is not x:
            np.testing.assert_array_equal(i, np.broadcast_to(np.arange(m, dtype=np.int32), i.shape))
        cur = np.take_along_axis(cur, i[..., None].astype(np.int64), axis=1)
    y = x.copy()
    y[:, 2048:] = y[:, :2048]                     # every point twice: the nested run must NOT be 0..m-1 everywhere
    i1 = cpu_ops.fps(y, 4096)
    lvl1 = np.take_along_axis(y, i1[..., None].astype(np.int64), axis=1)
    i2 = cpu_ops.fps(lvl1, 4096)
    assert (i2 != np.arange(4096)).any()


def test_lazy_log_is_a_full_mapping():
    """log_dict contract of PW/pwclo_net.py:186-193: a reader may use [], get, items, values, iteration, len, dict(), **"""
    from pwclonet_pylidarslam_b200.pwclonet.pwclo_net import LazyLog
    calls = []
    log = LazyLog({"embedding_mask": lambda: calls.append("m") or 1, "point_cloud": lambda: calls.append("p") or 2})
    assert len(log) == 2 and list(log) == ["embedding_mask", "point_cloud"] and "point_cloud" in log and calls == []
    assert log.get("embedding_mask") == 1 and log.get("nope", 7) == 7 and calls == ["m"]
    assert dict(log) == {"embedding_mask": 1, "point_cloud": 2} and calls == ["m", "p"]
    assert sorted(log.items()) == [("embedding_mask", 1), ("point_cloud", 2)] and list(log.values()) == [1, 2]
    assert (lambda **kw: kw)(**log) == {"embedding_mask": 1, "point_cloud": 2} and calls == ["m", "p"]


def test_knn_matches_torch_cpu_formulation():
    """the oracle's order-0 squared sums are bit-identical to the reference's torch expression on CPU;
    torch's CPU sqrt (MKL VML) is 1 ulp low on ~0.6 % of inputs, the oracle (like torch CUDA, which
    is what the reference actually runs on) uses the correctly rounded sqrt -- see DESIGN.md."""
    rng = np.random.default_rng(1)
    xyz = (rng.standard_normal((2, 700, 3)) * 9).astype(np.float32)
    q = xyz[:, :90] + 0.001
    t_xyz, t_q = torch.from_numpy(xyz), torch.from_numpy(q)
    diff = t_q.unsqueeze(2).repeat(1, 1, 700, 1) - t_xyz.unsqueeze(1).repeat(1, 90, 1, 1)
    ssum = torch.sum(diff ** 2, dim=-1) + 1e-8
    mine = cpu_ops.knn_distances(xyz, q, 0)
    np.testing.assert_array_equal(mine, np.sqrt(ssum.numpy()))           # IEEE sqrt of torch's own sum: exact
    dist = torch.sqrt(ssum).numpy()
    ulp = np.abs(mine.view(np.int32) - dist.view(np.int32))
    assert ulp.max() <= 1 and (ulp == 0).mean() > 0.98
    i = torch.topk(torch.from_numpy(mine), 16, largest=False, dim=-1)[1].numpy()
    oi = cpu_ops.knn(xyz, q, 16, 0)
    assert (np.sort(oi, -1) == np.sort(i, -1)).all()                     # same neighbour sets
    assert (oi == i).mean() > 0.999                                      # order differs only inside exact ties


@pytest.mark.parametrize("tag", ["forward_b2_n8192_w1", "forward_b1_n8192_w2"])
def test_port_reproduces_reference_golden(tag):
    """the travelling restatement (oracle/pwclo_port.py) against vectors produced by the UNMODIFIED reference"""
    g, x1, x2, wseed = C.load_golden(tag)
    w = C.weights_for(C.model_shapes(), wseed, g)
    port = Port(w)
    with torch.no_grad():
        pose, log = port.forward(x1, x2)
    te, re_ = C.pose_errors(pose.numpy(), g["pose"])
    assert te <= 1e-5 and re_ <= 1e-5, (te, re_)
    assert len(port.fps_log) == 9 and len(port.knn_log) == 23
    for i, f in enumerate(port.fps_log):
        np.testing.assert_array_equal(f.numpy(), g[f"fps_{i}"])
    for i, k in enumerate(port.knn_log):
        bad, bad_rows = C.knn_rows_match(k.numpy(), g, i)
        assert bad == 0 and bad_rows == 0, (i, bad, bad_rows)
    for l in range(1, 5):
        for fr in (1, 2):
            f = port.trace[f"f{fr}.psa{l}.feats"].numpy()
            assert C.rel_err(f[:, :, ::max(1, f.shape[2] // 64)], g[f"psa{l}_f{fr}_feats"]) <= 1e-5
    assert C.rel_err(port.trace["cv3.out"].numpy()[:, :, ::4], g["cv3_out"]) <= 1e-5
    for l in (3, 2, 1):
        e = port.trace[f"pwr{l}.emb"].numpy()
        assert C.rel_err(e[:, :, ::max(1, e.shape[2] // 64)], g[f"pwr{l}_emb"]) <= 1e-5
    assert C.rel_err(log["embedding_mask"].numpy()[:, ::16], g["log_embedding_mask"]) <= 1e-5


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted")
def test_port_equals_unmodified_reference_live():
    g, x1, x2, wseed = C.load_golden("forward_b1_n8192_w2")
    net = ref_shim.load_reference()
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    assert shapes == C.model_shapes()                       # our module has the reference's checkpoint layout
    w = C.weights_for(shapes, wseed, g)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    with torch.no_grad():
        ref_pose, _ = net(torch.from_numpy(x1), None, torch.from_numpy(x2), None)
        pose, _ = Port(w).forward(x1, x2)
    np.testing.assert_array_equal(ref_pose.numpy(), g["pose"])
    np.testing.assert_allclose(pose.numpy(), ref_pose.numpy(), rtol=0, atol=1e-6)


def test_state_dict_layout_and_reference_checkpoint_loading():
    from pwclonet_pylidarslam_b200.pwclonet import PWCLONet
    net = PWCLONet({"device": "cpu"})
    sd = net.state_dict()
    assert len(sd) == 510
    assert sum(v.numel() for k, v in sd.items() if k.endswith(("conv.weight", "conv.bias", "bn.weight", "bn.bias"))) == 775068
    assert sd["psa_1.mlp_module.layer0.conv.weight"].shape == (8, 6, 1, 1)
    assert sd["pose_warp_refinement_1.cost_volume.mlp3_convs.layer0.conv.weight"].shape == (128, 144, 1, 1)
    assert "pose_warp_refinement_1.flow_predictor_mask.mlp_convs.layer0.conv.weight" not in sd
    assert sd["pose_calculator_4.conv1d_q_t.conv.weight"].shape == (256, 64, 1)
    # a checkpoint saved by the reference trainer nests the module under `pwclonet.` (prediction_modules.py:127)
    w = C.weights_for({k: tuple(v.shape) for k, v in sd.items()}, 3)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})


def test_c_abi_exports_every_declared_symbol():
    from pwclonet_pylidarslam_b200 import _lib
    header = open(os.path.join(ROOT, "include", "pwclo_b200.h")).read()
    declared = set(re.findall(r"\b(pwclo_[a-z0-9_]+)\s*\(", header))
    assert declared, "header parse failed"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/pwclo_b200.h but not exported"
    assert declared - {"pwclo_version", "pwclo_error_string", "pwclo_knn_workspace_bytes", "pwclo_prepare_scans_workspace_bytes",
                       "pwclo_bn_relu_workspace_bytes", "pwclo_conv1x1_wgrad_workspace_bytes"} == set(_lib.SIGNATURES), "ctypes table out of sync"
    assert b"sm_100a" in _lib.lib().pwclo_version()


def test_product_never_imports_oracle():
    """the product package must not route through the oracle (or any CPU fallback)"""
    pkg = os.path.join(ROOT, "pwclonet_pylidarslam_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("oracle-of-record", ""), f"{f} mentions the oracle"


def test_ext_refuses_cpu_tensors():
    from pwclonet_pylidarslam_b200 import _ext
    with pytest.raises(RuntimeError):
        _ext.furthest_point_sampling(torch.zeros(1, 8, 3), 2)
    with pytest.raises(RuntimeError):
        _ext.knn(torch.zeros(1, 8, 3), torch.zeros(1, 2, 3), 2)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted")
def test_small_helpers_equal_reference():
    """switch_quat and set_bn_momentum_default behave as the reference's (module surface parity)"""
    ref_shim.load_reference()
    mods = ref_shim.reference_modules()
    from pwclonet_pylidarslam_b200 import pytorch_utils as ours_pt
    from pwclonet_pylidarslam_b200.pwclonet import PWCLO_utils as ours_u
    rng = np.random.default_rng(0)
    for shape in ((4,), (5, 4)):
        q = rng.standard_normal(shape).astype(np.float32)
        for last in (True, False):
            np.testing.assert_array_equal(ours_u.switch_quat(q, last), mods["PWCLO_utils"].switch_quat(q, last))
    with pytest.raises(RuntimeError):
        ours_u.switch_quat(np.zeros((2, 3, 4)))
    a, b = torch.nn.Sequential(torch.nn.BatchNorm2d(3), torch.nn.Conv2d(3, 3, 1)), torch.nn.Sequential(torch.nn.BatchNorm2d(3))
    a.apply(ours_pt.set_bn_momentum_default(0.37))
    b.apply(mods["pytorch_utils"].set_bn_momentum_default(0.37))
    assert a[0].momentum == b[0].momentum == 0.37
