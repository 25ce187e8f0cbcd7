"""Full-forward parity on the GPU against golden vectors produced by the
UNMODIFIED reference (tests/golden, oracle/make_golden.py) and against the travelling oracle port.
Tolerances are the ones BASELINE.json states: FPS / kNN indices bit-exact (kNN: modulo exact-tie
order), features 1e-4 relative, pose 1e-4 m / 1e-5 rad."""
import numpy as np
import pytest
import torch

from tests import _common as C

pytestmark = pytest.mark.gpu


def _net(cuda, wseed, g, fused):
    from pwclonet_pylidarslam_b200.pwclonet import PWCLONet
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net = PWCLONet({"device": "cuda:0", "use_fused": fused})
    w = C.weights_for({k: tuple(v.shape) for k, v in net.state_dict().items()}, wseed, g)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    return net.to(cuda).eval(), w


@pytest.mark.parametrize("tag", ["forward_b2_n8192_w1", "forward_b1_n8192_w2"])
@pytest.mark.parametrize("fused", [False, True])
def test_forward_matches_reference_golden(cuda, tag, fused):
    from pwclonet_pylidarslam_b200 import _ext
    g, x1, x2, wseed = C.load_golden(tag)
    net, _ = _net(cuda, wseed, g, fused)
    old = _ext.KNN_SUM_ORDER
    _ext.KNN_SUM_ORDER = 0      # the golden was produced by the reference on CPU: torch-CPU summation order
    try:
        with torch.no_grad():
            pose, log = net(torch.from_numpy(x1).to(cuda), None, torch.from_numpy(x2).to(cuda), None)
    finally:
        _ext.KNN_SUM_ORDER = old
    te, re_ = C.pose_errors(pose.cpu().numpy(), g["pose"])
    print(f"{tag} fused={fused}: translation err {te:.3e} m, rotation err {re_:.3e} rad")
    assert te <= C.TOL_TRANSLATION_M and re_ <= C.TOL_ROTATION_RAD
    assert C.rel_err(log["embedding_mask"].numpy()[:, ::16], g["log_embedding_mask"]) <= C.TOL_FEATURE_REL
    assert log["point_cloud"].shape == (x1.shape[0], 2048, 3)


def test_forward_gpu_order_vs_port(cuda):
    """default (torch-CUDA) kNN summation order against the oracle port configured the same way"""
    from oracle.pwclo_port import Port
    g, x1, x2, wseed = C.load_golden("forward_b1_n8192_w2")
    for fused in (False, True):
        net, w = _net(cuda, wseed, g, fused)
        with torch.no_grad():
            pose, _ = net(torch.from_numpy(x1).to(cuda), None, torch.from_numpy(x2).to(cuda), None)
            want, _ = Port(w, sum_order=1).forward(x1, x2)
        te, re_ = C.pose_errors(pose.cpu().numpy(), want.numpy())
        assert te <= C.TOL_TRANSLATION_M and re_ <= C.TOL_ROTATION_RAD, (fused, te, re_)


def test_pose_pipeline_equals_direct_forward(cuda):
    """sharding.PosePipeline (copy of batch i+1 overlapped with the forward of batch i) returns, in order, exactly
    what the direct calls return"""
    from pwclonet_pylidarslam_b200 import synthetic as syn
    from pwclonet_pylidarslam_b200.pwclonet import PWCLONet
    from pwclonet_pylidarslam_b200.sharding import PosePipeline
    net = PWCLONet({"device": "cuda:0"}).to(cuda).eval()
    batches = []
    for i in range(5):
        x1, x2, _ = syn.make_batch(300 + 2 * i, 2, 4096)
        batches.append((torch.from_numpy(x1).pin_memory(), torch.from_numpy(x2).pin_memory()))
    direct = []
    with torch.no_grad():
        for h1, h2 in batches:
            direct.append(net(h1.to(cuda), None, h2.to(cuda), None)[0].cpu())
    pipe = PosePipeline(net, 2, 4096)
    got = [p.clone() for p in pipe.run(iter(batches))]
    assert len(got) == 5
    for a, b in zip(got, direct):
        assert torch.equal(a, b)


def _rand_net(cuda, seed, **cfg):
    from pwclonet_pylidarslam_b200.pwclonet import PWCLONet
    net = PWCLONet({"device": "cuda:0", **cfg})
    w = C.weights_for({k: tuple(v.shape) for k, v in net.state_dict().items()}, seed)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    return net.to(cuda).eval(), w


def test_graphed_forward_equals_eager_forward(cuda):
    """the whole forward replayed from one CUDA graph (FusedPWCLONet.forward_graphed) returns bit for bit what the
    launch-by-launch forward returns, for several shapes and for fresh inputs on every replay; results of earlier
    calls stay valid after later calls (outputs are cloned out of the graph's static buffers)"""
    from pwclonet_pylidarslam_b200 import synthetic as syn
    net, _ = _rand_net(cuda, 4)
    eng = net.fused_engine()
    kept = []
    for B, N, first in ((1, 8192, 500), (3, 8192, 510), (1, 8192, 520), (2, 4096, 530), (3, 8192, 540)):
        x1, x2, _ = syn.make_batch(first, B, N)
        a, b = torch.from_numpy(x1).to(cuda), torch.from_numpy(x2).to(cuda)
        with torch.no_grad():
            pe, me, xe = eng.forward(a, b)
            pg, log = net(a, None, b, None)                  # the module's default eval path: graphed
        assert eng._graphs.get((B, N)) not in (None, False), "graph capture did not happen"
        assert torch.equal(pe, pg)
        assert torch.equal(log["point_cloud"], xe.cpu())
        kept.append((pg, pe.clone()))
    for pg, pe in kept:
        assert torch.equal(pg, pe)
    assert eng.launches > 0


def test_eval_forward_without_no_grad_takes_the_fused_engine(cuda):
    """the reference's test path calls the model in eval() with autograd enabled (train.py:348-359, :806-808):
    it must get the fused engine, not silently the composed torch path; `eval_autograd=True` keeps the graph"""
    from pwclonet_pylidarslam_b200 import synthetic as syn
    net, _ = _rand_net(cuda, 6)
    x1, x2, _ = syn.make_batch(700, 1, 8192)
    a, b = torch.from_numpy(x1).to(cuda), torch.from_numpy(x2).to(cuda)
    n0 = net.fused_engine().launches
    pose, _ = net(a, None, b, None)                          # grad mode is ON here
    assert net.fused_engine().launches > n0 and not pose.requires_grad
    with torch.no_grad():
        want, _ = net(a, None, b, None)
    assert torch.equal(pose, want)
    net2, _ = _rand_net(cuda, 6, eval_autograd=True)
    pose2, _ = net2(a, None, b, None)
    assert pose2.requires_grad and net2._fused is None
    te, re_ = C.pose_errors(pose2.detach().cpu().numpy(), want.cpu().numpy())
    assert te <= C.TOL_TRANSLATION_M and re_ <= C.TOL_ROTATION_RAD


def test_fused_cache_follows_the_weights(cuda):
    """eval forward -> new weights through a PARENT module's load_state_dict (the trainer's checkpoint
    path), through .to(), through an in-place torch update -> the next eval forward must use the new weights"""
    from pwclonet_pylidarslam_b200 import synthetic as syn
    from pwclonet_pylidarslam_b200.training import _PWCLONetPredictionModule
    x1, x2, _ = syn.make_batch(710, 1, 8192)
    a, b = torch.from_numpy(x1).to(cuda), torch.from_numpy(x2).to(cuda)
    parent = _PWCLONetPredictionModule({"device": "cuda:0"}).to(cuda).eval()
    net = parent.pwclonet
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    w1, w2 = C.weights_for(shapes, 11), C.weights_for(shapes, 12)
    fresh = {}
    for tag, w in (("w1", w1), ("w2", w2)):
        f, _ = _rand_net(cuda, 11 if tag == "w1" else 12)
        with torch.no_grad():
            fresh[tag] = f(a, None, b, None)[0]
    assert not torch.equal(fresh["w1"], fresh["w2"])
    parent.load_state_dict({"pwclonet." + k: torch.from_numpy(v) for k, v in w1.items()})
    with torch.no_grad():
        assert torch.equal(net(a, None, b, None)[0], fresh["w1"])
        parent.load_state_dict({"pwclonet." + k: torch.from_numpy(v) for k, v in w2.items()})      # parent recursion
        assert torch.equal(net(a, None, b, None)[0], fresh["w2"])
        for k, p in net.state_dict().items():                                                    # in-place torch update
            p.copy_(torch.from_numpy(w1[k]))
        assert torch.equal(net(a, None, b, None)[0], fresh["w1"])
        parent.float().to(cuda)                                                                  # _apply
        assert net._fused is None


@pytest.mark.parametrize("n", [2, 6])
def test_forward_streams_in_flight_equal_direct(cuda, n):
    """sharding.ForwardStreams: forwards submitted back to back on n compute streams (own graph + static buffers per
    stream; six is the default policy) return exactly what one-at-a-time calls return, for inputs that change on every
    submission and with every stream's graph replayed at least twice"""
    from pwclonet_pylidarslam_b200 import synthetic as syn
    from pwclonet_pylidarslam_b200.sharding import ForwardStreams, auto_compute_streams
    net, _ = _rand_net(cuda, 8)
    ins = []
    for i in range(7):
        x1, x2, _ = syn.make_batch(600 + 3 * i, 3, 8192)
        ins.append((torch.from_numpy(x1).to(cuda), torch.from_numpy(x2).to(cuda)))
    with torch.no_grad():
        direct = [net(a, None, b, None)[0].clone() for a, b in ins]
    fwd = ForwardStreams(net, n)
    order = [i % len(ins) for i in range(2 * n + 1)]
    got = [fwd.submit(*ins[i]) for i in order]
    fwd.join()
    torch.cuda.synchronize()
    assert len({id(s) for s in fwd.streams}) == n and fwd.count == len(order)
    for (pose, _ev), i in zip(got, order):
        assert torch.equal(pose, direct[i])
    assert auto_compute_streams(3) >= 1


@pytest.mark.parametrize("B,N", [(1, 2048), (2, 4096), (1, 16384)])
def test_forward_other_cloud_sizes_vs_port(cuda, B, N):
    """the level sizes are fixed (2048 / 1024 / 256 / 64), the input size is not: N = 2048 (level 1 keeps every point),
    4096, and the 16 384-point clouds of the training configuration (cluster FPS, brute-force level-1 search) -- fused
    forward (graph replay) against the travelling port of the reference, same kNN summation order"""
    from oracle.pwclo_port import Port
    from pwclonet_pylidarslam_b200 import synthetic as syn
    net, w = _rand_net(cuda, 21 + B)
    x1, x2, _ = syn.make_batch(640 + N % 97, B, N)
    with torch.no_grad():
        pose, log = net(torch.from_numpy(x1).to(cuda), None, torch.from_numpy(x2).to(cuda), None)
        want, wlog = Port(w, sum_order=1).forward(x1, x2)
    te, re_ = C.pose_errors(pose.cpu().numpy(), want.numpy())
    print(f"B={B} N={N}: translation {te:.2e} m, rotation {re_:.2e} rad")
    assert te <= C.TOL_TRANSLATION_M and re_ <= C.TOL_ROTATION_RAD
    np.testing.assert_array_equal(log["point_cloud"].numpy(), wlog["point_cloud"].numpy())
