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
