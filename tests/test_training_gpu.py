"""Training-mode path: the composed module runs forward + backward on the sm_100a operators
(grouping scatter-add gradients, non-differentiable FPS / kNN indices), gradients are finite and agree
with a pure-PyTorch evaluation of the same graph (index_select based grouping)."""
import numpy as np
import pytest
import torch

from tests import _common as C

pytestmark = pytest.mark.gpu


def test_grouping_backward_matches_torch(cuda):
    from pwclonet_pylidarslam_b200 import pointnet2_utils as pu
    g = torch.Generator(device=cuda).manual_seed(0)
    f = torch.randn(2, 16, 500, device=cuda, generator=g, requires_grad=True)
    idx = torch.randint(0, 500, (2, 64, 8), device=cuda, dtype=torch.int32, generator=g)
    w = torch.randn(2, 16, 64, 8, device=cuda, generator=g)
    (pu.grouping_operation(f, idx) * w).sum().backward()
    got = f.grad.clone()
    f.grad = None
    ref = torch.gather(f.unsqueeze(2).expand(-1, -1, 64, -1), 3, idx.long().unsqueeze(1).expand(-1, 16, -1, -1))
    (ref * w).sum().backward()
    torch.testing.assert_close(got, f.grad, rtol=1e-5, atol=1e-5)


def test_train_step_forward_backward(cuda):
    from pwclonet_pylidarslam_b200.pwclonet import PWCLONet
    g, x1, x2, wseed = C.load_golden("forward_b1_n8192_w2")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net = PWCLONet({"device": "cuda:0"})
    w = C.weights_for({k: tuple(v.shape) for k, v in net.state_dict().items()}, wseed, g)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    net = net.to(cuda).train()
    torch.manual_seed(0)
    pose, _ = net(torch.from_numpy(x1).to(cuda), None, torch.from_numpy(x2).to(cuda), None)
    assert pose.shape == (1, 4, 7) and pose.requires_grad
    # loss in the spirit of slam/training/loss_modules.py:424-544 (L2 on t, L2 on q, level weights)
    gt = torch.tensor([[0.1, 0.0, 1.0, 1.0, 0.0, 0.0, 0.0]], device=cuda)
    wts = torch.tensor([0.2, 0.4, 0.8, 1.6], device=cuda)
    loss = (wts * ((pose[..., :3] - gt[:, None, :3]).norm(dim=-1) + (pose[..., 3:] - gt[:, None, 3:]).norm(dim=-1))).sum()
    loss.backward()
    grads = [p.grad for p in net.parameters()]
    assert all(g_ is not None and torch.isfinite(g_).all() for g_ in grads)
    total = float(sum(float(g_.abs().sum()) for g_ in grads))
    assert total > 0
    # level-1 set conv only has weight gradients (its input is raw xyz): they must be non-zero
    assert float(net.psa_1.mlp_module.layer0.conv.weight.grad.abs().sum()) > 0
