"""Training-mode path: the composed module runs forward + backward on the sm_100a operators
(grouping scatter-add gradients, non-differentiable FPS / kNN indices); the whole step is pinned against a golden
produced by the UNMODIFIED reference (model in train(), reference loss, torch autograd), the operators against
pure-PyTorch evaluations of the same graph."""
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


def _reference_train_step_on_gpu(cuda, ref_ext, w, x1, x2, gt, g):
    """the UNMODIFIED reference model (staged python + its own CUDA extension) doing the same training step on this GPU:
    its distance to the CPU golden is the device-to-device noise floor of this computation"""
    from oracle import ref_shim
    if ref_ext is None or not ref_shim.available():
        return None
    ref = ref_shim.load_reference(ext_module=ref_ext, device="cuda:0").to(cuda)
    ref.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    ref.train()
    from pwclonet_pylidarslam_b200 import training as T
    loss_mod = T._PWCLONetLossModule(T.PWCLONetLossConfig()).to(cuda)    # pinned against the reference loss (loss_kat.npz)
    pose, _ = ref(torch.from_numpy(x1).to(cuda), None, torch.from_numpy(x2).to(cuda), None)
    loss, _ = loss_mod(pose, torch.from_numpy(gt).to(cuda))
    loss.backward()
    params = dict(ref.named_parameters())
    errs = [C.rel_err(params[str(n)].grad.cpu().numpy(), g[f"grad_{i}"]) for i, n in enumerate(g["grad_names"])]
    norms = np.asarray([float(p.grad.double().norm()) for p in params.values()])
    rel = np.abs(norms - g["grad_norm_all"]) / np.maximum(g["grad_norm_all"], 1e-30)
    return max(errs), float(rel.max())


def test_training_step_matches_unmodified_reference_golden(cuda, ref_ext, monkeypatch):
    """One whole training step against the UNMODIFIED reference (tests/golden/train_step_b2_n2048.npz, written by
    oracle/make_golden_train_step.py: reference model in train() on CPU, reference loss module, torch autograd):
    train-mode BatchNorm (batch statistics + running-statistics update), every backward on the path (grouping
    scatter-add, pose warp, BN + ReLU, few-channel convolutions, loss) -- loss, pose, the gradient of 20 named tensors
    covering every kind of layer, the gradient norm of ALL 318 parameter tensors, and updated running statistics.
    fp32 everywhere (TF32 off), torch-CPU kNN summation order, dropout = identity on both sides."""
    import os
    import torch.nn.functional as F_
    from pwclonet_pylidarslam_b200 import _ext, synthetic as syn
    from pwclonet_pylidarslam_b200 import training as T
    from pwclonet_pylidarslam_b200.pwclonet import PWCLONet
    monkeypatch.setattr(F_, "dropout", lambda x, p=0.5, training=True, inplace=False: x)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = dict(np.load(os.path.join(C.GOLD_DIR, "train_step_b2_n2048.npz")))
    first, pairs, points, wseed = [int(v) for v in g["meta"]]
    x1, x2, gt = syn.make_batch(first, pairs, points)
    assert C.sha(x1, x2, gt) == str(g["input_sha"]), "synthetic generator drifted from the golden inputs"
    net = PWCLONet({"device": "cuda:0"})
    w = syn.make_state_dict({k: tuple(v.shape) for k, v in net.state_dict().items()}, seed=wseed)
    assert C.sha(*[w[k] for k in sorted(w)]) == str(g["weights_sha"])
    net.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    net = net.to(cuda).train()
    loss_mod = T._PWCLONetLossModule(T.PWCLONetLossConfig()).to(cuda)
    monkeypatch.setattr(_ext, "KNN_SUM_ORDER", 0)       # the golden ran torch's CPU reduction order
    nbt = {k: int(v) for k, v in net.state_dict().items() if k.endswith("num_batches_tracked")}
    pose, _ = net(torch.from_numpy(x1).to(cuda), None, torch.from_numpy(x2).to(cuda), None)
    # nn.BatchNorm bookkeeping (applied in one multi-tensor add at the end of the forward): the siamese pyramid layers ran
    # once per frame, everything else once
    after = {k: int(v) for k, v in net.state_dict().items() if k.endswith("num_batches_tracked")}
    assert all(after[k] - nbt[k] == (2 if k.startswith("psa_") else 1) for k in nbt), \
        {k: after[k] - nbt[k] for k in nbt if after[k] - nbt[k] != (2 if k.startswith("psa_") else 1)}
    loss, _ = loss_mod(pose, torch.from_numpy(gt).to(cuda))
    loss.backward()
    te, re_ = C.pose_errors(pose.detach().cpu().numpy(), g["pose"])
    loss_v = float(loss.detach())
    params = dict(net.named_parameters())
    errs = {str(n): C.rel_err(params[str(n)].grad.cpu().numpy(), g[f"grad_{i}"]) for i, n in enumerate(g["grad_names"])}
    norms = np.asarray([float(p.grad.double().norm()) for p in params.values()])
    rel = np.abs(norms - g["grad_norm_all"]) / np.maximum(g["grad_norm_all"], 1e-30)
    sd = net.state_dict()
    stat = {str(n): max(C.rel_err(sd[str(n) + ".running_mean"].cpu().numpy(), g[f"mean_{i}"]),
                        C.rel_err(sd[str(n) + ".running_var"].cpu().numpy(), g[f"var_{i}"])) for i, n in enumerate(g["stat_names"])}
    worst = max(errs, key=errs.get)
    print(f"train-mode forward vs reference: translation {te:.2e} m, rotation {re_:.2e} rad, loss {loss_v:.6f} / {float(g['loss']):.6f}")
    print(f"worst relative gradient error over the {len(errs)} tensors: {errs[worst]:.2e} ({worst}); worst gradient-norm error over "
          f"all {len(norms)} tensors {rel.max():.2e}; worst running-statistics error {max(stat.values()):.2e}")
    # Train-mode BatchNorm divides by the statistics of a 2-pair batch and its backward subtracts batch means, which
    # amplifies the fp32 summation-order differences between torch-CPU and any GPU implementation.  The pose tolerance of the
    # INFERENCE path (1e-4 m / 1e-5 rad) is widened to 1e-4 m / 1e-4 rad here, and the gradient tolerance is anchored on the
    # noise floor measured in the same run: the UNMODIFIED reference executing this very step on this GPU against its
    # own CPU golden.  Ours must be within 1e-3 or 3x that floor, whichever is larger, and never beyond 1e-2.
    noise = _reference_train_step_on_gpu(cuda, ref_ext, w, x1, x2, gt, g)
    print(f"reference-on-GPU vs its own CPU golden (noise floor): worst gradient error / worst gradient-norm error = {noise}")
    tol_g, tol_n = (1e-3, 1e-3) if noise is None else (min(1e-2, max(1e-3, 3 * noise[0])), min(2e-2, max(1e-3, 3 * noise[1])))
    assert te <= 1e-4 and re_ <= 1e-4
    assert abs(loss_v - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert errs[worst] <= tol_g, (tol_g, errs)
    assert rel.max() <= tol_n, (tol_n, float(rel.max()))
    np.testing.assert_allclose(loss_mod.exp_weighting.s_param.grad.cpu().numpy(), g["grad_s"], rtol=1e-4)
    assert max(stat.values()) <= 1e-4, stat


# ---- training-side rows (F15 / N1): loss + gradient kernel, flat Adam, trainer, checkpoints -------------
def _loss_cases():
    import os
    g = dict(np.load(os.path.join(C.GOLD_DIR, "loss_kat.npz")))
    for i in range(int(g["n_cases"])):
        yield i, {k[len(f"c{i}_"):]: v for k, v in g.items() if k.startswith(f"c{i}_")}


def test_pose_loss_kernel_matches_reference_golden(cuda):
    """value and gradient of pwclo_pose_loss against the unmodified reference loss module (fixture made by
    oracle/make_golden_train.py).  fp32 with a different summation order than torch.sum / torch.mean: 2e-6
    relative on the terms; gradients 1e-5 relative to the largest entry for O(1) errors, 1e-4 for the "near"
    cases (pose errors ~1e-3, as for a trained network): there q/|q| - q_gt cancels three digits, so a 1-ulp
    difference in |q| (summation order of four squares) moves the gradient by ~5e-5 relative in ANY fp32
    evaluation, the reference's included."""
    from pwclonet_pylidarslam_b200 import training as T
    for i, c in _loss_cases():
        with_exp = bool(c["with_exp"])
        cfg = T.PWCLONetLossConfig(with_exp_weights=with_exp, init_weights=[float(v) for v in c["s"]],
                                   loss_weights=[float(v) for v in c["s"]])
        mod = T._PWCLONetLossModule(cfg).to(cuda)
        p = torch.tensor(c["pred"], device=cuda, requires_grad=True)
        loss, log = mod(p, torch.tensor(c["gt"], device=cuda))
        loss.backward()
        keys = (["loss"] + [f"loss_l{j}" for j in range(1, 5)] + [f"loss_rot_l{j}" for j in range(1, 5)]
                + [f"loss_trans_l{j}" for j in range(1, 5)])
        got = np.array([float(log[k]) for k in keys], np.float32)
        np.testing.assert_allclose(got, c["terms"], rtol=2e-6, atol=1e-7, err_msg=f"case {i}")
        gp = p.grad.cpu().numpy()
        near = float(np.abs(c["pred"] - c["gt"][:, None, :]).max()) < 0.05
        tol = 1e-4 if near else 1e-5
        assert np.abs(gp - c["grad_pred"]).max() <= tol * np.abs(c["grad_pred"]).max() + 1e-9, f"case {i}"
        if with_exp:
            gs = mod.exp_weighting.s_param.grad.cpu().numpy()
            np.testing.assert_allclose(gs, c["grad_s"], rtol=1e-5, atol=1e-6, err_msg=f"case {i}")


def test_pose_loss_kernel_matches_oracle_large_batch(cuda):
    from oracle import train_port
    from pwclonet_pylidarslam_b200 import training as T
    rng = np.random.default_rng(11)
    pred = rng.standard_normal((1000, 4, 7)).astype(np.float32)
    gt = rng.standard_normal((1000, 7)).astype(np.float32)
    o = train_port.pose_loss(pred, gt, [0.1, -2.0])
    mod = T._PWCLONetLossModule(T.PWCLONetLossConfig(init_weights=[0.1, -2.0])).to(cuda)
    p = torch.tensor(pred, device=cuda, requires_grad=True)
    loss, _ = mod(p, torch.tensor(gt, device=cuda))
    (2.0 * loss).backward()                          # upstream gradient is honoured
    assert abs(float(loss) - o["loss"]) <= 2e-6 * abs(o["loss"])
    assert np.abs(p.grad.cpu().numpy() - 2.0 * o["grad_pred"]).max() <= 1e-5 * np.abs(2.0 * o["grad_pred"]).max()


def test_flat_adam_matches_torch_adam(cuda):
    """pwclo_adam_step over the flat arena == torch.optim.Adam on the same tensors (fp32 rounding only)"""
    from pwclonet_pylidarslam_b200 import training as T
    torch.manual_seed(0)
    a = torch.nn.Sequential(torch.nn.Linear(33, 65), torch.nn.ReLU(), torch.nn.Linear(65, 7)).to(cuda)
    b = torch.nn.Sequential(torch.nn.Linear(33, 65), torch.nn.ReLU(), torch.nn.Linear(65, 7)).to(cuda)
    b.load_state_dict(a.state_dict())
    arena = T.FlatArena([a])
    ours = T.FlatAdam(arena, lr=1e-3, weight_decay=1e-3)
    ref = torch.optim.Adam(b.parameters(), lr=1e-3, betas=(0.9, 0.999), weight_decay=1e-3, foreach=False)
    for step in range(6):
        x = torch.randn(16, 33, device=cuda)
        ours.zero_grad()
        ref.zero_grad()
        a(x).square().mean().backward()
        b(x).square().mean().backward()
        ours.step()
        ref.step()
    for pa, pb in zip(a.parameters(), b.parameters()):
        torch.testing.assert_close(pa, pb, rtol=1e-5, atol=1e-6)
    # the optimiser state moves to torch.optim.Adam and back (checkpoint compatibility, trainer.py:884)
    ref2 = torch.optim.Adam(b.parameters(), lr=1e-3, weight_decay=1e-3)
    ref2.load_state_dict(ours.state_dict())
    assert int(ref2.state_dict()["state"][0]["step"]) == 6
    again = T.FlatAdam(T.FlatArena([b]), lr=5e-4)
    again.load_state_dict(ref.state_dict())
    assert again.steps == 6 and abs(again.lr - 1e-3) < 1e-12
    torch.testing.assert_close(again.exp_avg, ours.exp_avg, rtol=1e-4, atol=1e-7)


def test_trainer_step_and_checkpoint_roundtrip(cuda, tmp_path):
    """one PWCLONetTrainer: a few steps on a fixed batch reduce the loss; checkpoint has the reference's
    keys (trainer.py:882-907) and restores parameters, optimiser state and schedules exactly."""
    from pwclonet_pylidarslam_b200 import synthetic as syn
    from pwclonet_pylidarslam_b200 import training as T
    torch.manual_seed(0)
    cfg = T.PWCLONetTrainerConfig(num_points=4096, num_epochs=10, optimizer_weight_decay=0.0)
    tr = T.PWCLONetTrainer(cfg)
    assert tr.arena.count == 775070                        # 775 068 + 2
    x1, x2, gt = syn.make_batch(700, 2, 4096)
    batch = [torch.from_numpy(np.ascontiguousarray(x1.transpose(0, 2, 1))).to(cuda),
             torch.from_numpy(np.ascontiguousarray(x2.transpose(0, 2, 1))).to(cuda),
             torch.from_numpy(gt[:, 3:]).to(cuda), torch.from_numpy(gt[:, :3]).to(cuda)]
    losses = []
    for _ in range(6):
        loss, log, pred = tr.train_step(batch)
        losses.append(float(loss))
        assert pred.shape == (2, 4, 7)
    assert all(np.isfinite(losses)) and min(losses[1:]) < losses[0], losses
    tr.end_epoch()
    ck = str(tmp_path / "0.ckp")
    tr.save_checkpoint(ck)
    sd = torch.load(ck, weights_only=False)
    assert {"optimizer", "loss_module", "prediction_module", "num_train_epochs", "train_iter", "eval_iter", "best"} <= set(sd)
    assert len(sd["prediction_module"]) == 510 and all(k.startswith("pwclonet.") for k in sd["prediction_module"])
    tr2 = T.PWCLONetTrainer(cfg)
    tr2.load_checkpoint(ck)
    assert torch.equal(tr2.arena.param, tr.arena.param) and torch.equal(tr2._optimizer.exp_avg_sq, tr._optimizer.exp_avg_sq)
    assert tr2._optimizer.steps == 6 and tr2.num_epochs == 1 and tr2._optimizer.lr == tr._optimizer.lr
    # both continue identically (dropout seeds aligned)
    torch.manual_seed(5)
    l1, _, _ = tr.train_step(batch)
    torch.manual_seed(5)
    l2, _, _ = tr2.train_step(batch)
    assert abs(float(l1) - float(l2)) <= 1e-5 * abs(float(l1))


def test_graphed_step_equals_eager_step(cuda, monkeypatch):
    """the captured CUDA graphs (step + prefetched geometry) reproduce the eager step: from identical state, the same loss
    and the same updated parameters, step after step on alternating batches (the two dropouts of the pose head draw
    different masks eagerly and under replay, so they are switched off for this comparison)"""
    import torch.nn.functional as F_
    monkeypatch.setattr(F_, "dropout", lambda x, p=0.5, training=True, inplace=False: x)
    from pwclonet_pylidarslam_b200 import synthetic as syn
    from pwclonet_pylidarslam_b200 import training as T
    x1, x2, gt = syn.make_batch(710, 2, 4096)
    batch = [torch.from_numpy(np.ascontiguousarray(x1.transpose(0, 2, 1))).to(cuda),
             torch.from_numpy(np.ascontiguousarray(x2.transpose(0, 2, 1))).to(cuda),
             torch.from_numpy(gt[:, 3:]).to(cuda), torch.from_numpy(gt[:, :3]).to(cuda)]
    batch2 = [b.flip(0).contiguous() for b in batch]
    cfg = T.PWCLONetTrainerConfig(num_points=4096, num_epochs=10)
    torch.manual_seed(0)
    a = T.PWCLONetTrainer(cfg)
    b = T.PWCLONetTrainer(cfg)
    b.prediction_module_.load_state_dict(a.prediction_module_.state_dict())
    def sync_b_to_a():
        b.prediction_module_.load_state_dict(a.prediction_module_.state_dict())
        b.loss_module_.load_state_dict(a.loss_module_.state_dict())
        b._optimizer.load_state_dict(a._optimizer.state_dict())

    for i in range(6):
        a.train_step(batch if i % 2 == 0 else batch2)
    b.capture(batch, warmup=1)
    # Every comparison starts from identical state (a -> b): it checks ONE step (forward through the loss, update through
    # the parameters), not how fast two trajectories of a chaotic system drift apart under the run-to-run noise of the
    # scatter-add atomics.  Adam's normalised update may flip on gradient entries that are pure noise: a few lr at most.
    lr = a._optimizer.lr
    for i in range(4):
        bt, nxt = (batch, batch2) if i % 2 == 0 else (batch2, batch)
        sync_b_to_a()
        la = float(a.train_step(bt)[0])
        # steps 0-1 prefetch the next batch's geometry on the second stream, step 2 announces the WRONG next batch (the
        # prefetch must be discarded), step 3 announces none (geometry replayed in line)
        lb = float(b.train_step_graphed(bt, next_batch=(nxt if i < 2 else (bt if i == 2 else None)))[0])
        dp = (a.arena.param - b.arena.param).abs()
        print(f"step {i}: eager {la:.6f} graphed {lb:.6f} max |dparam| {float(dp.max()) / lr:.2f} lr, mean {float(dp.mean()) / lr:.4f} lr")
        assert abs(la - lb) <= 1e-5 * abs(la), (i, la, lb)
        assert float(dp.max()) <= 3.0 * lr and float(dp.mean()) <= 0.2 * lr, (i, float(dp.max()) / lr, float(dp.mean()) / lr)
    assert b._geo_graph is not None
    assert b._optimizer.steps == a._optimizer.steps == 10
    b._optimizer.lr = 5e-4                           # learning-rate changes reach the replayed graph
    p0 = b.arena.param.clone()
    b.train_step_graphed(batch)
    d1 = float((b.arena.param - p0).abs().max())
    assert 0 < d1 < 1e-2


@pytest.mark.parametrize("shape", [(4, 16, 512, 32), (16, 8, 2048, 32), (3, 64, 37, 5), (2, 128, 64), (8, 256, 1, 1), (2, 7, 1000)])
def test_fused_bn_relu_train_matches_torch(cuda, shape):
    """pwclo_bn_relu_train_fwd/_bwd against nn.BatchNorm (train mode, cuDNN / ATen) + ReLU on the same GPU: outputs,
    running statistics and all three gradients.  fp32 with double-accumulated statistics: 2e-5 relative to the
    largest entry (torch accumulates in fp32)."""
    from pwclonet_pylidarslam_b200.pytorch_utils import FusedBNReLUTrain
    g = torch.Generator(device=cuda).manual_seed(sum(shape))
    C = shape[1]
    x = (torch.randn(shape, device=cuda, generator=g) * 2.0 + 0.7).requires_grad_(True)
    w = (torch.rand(C, device=cuda, generator=g) + 0.5).requires_grad_(True)
    b = (torch.randn(C, device=cuda, generator=g) * 0.3).requires_grad_(True)
    up = torch.randn(shape, device=cuda, generator=g)
    rm0, rv0 = torch.randn(C, device=cuda, generator=g) * 0.1, torch.rand(C, device=cuda, generator=g) + 0.5
    rm, rv = rm0.clone(), rv0.clone()
    y = FusedBNReLUTrain.apply(x, w, b, rm, rv, 0.3, 1e-5)
    (y * up).sum().backward()
    got = [y.detach(), x.grad.clone(), w.grad.clone(), b.grad.clone(), rm, rv]
    x.grad = w.grad = b.grad = None
    rm2, rv2 = rm0.clone(), rv0.clone()
    yr = torch.relu(torch.nn.functional.batch_norm(x, rm2, rv2, w, b, True, 0.3, 1e-5))
    (yr * up).sum().backward()
    want = [yr.detach(), x.grad, w.grad, b.grad, rm2, rv2]
    for name, a, r in zip(("y", "dx", "dgamma", "dbeta", "running_mean", "running_var"), got, want):
        err = float((a - r).abs().max())
        assert err <= 2e-5 * float(r.abs().max()) + 1e-7, (name, shape, err, float(r.abs().max()))


@pytest.mark.parametrize("B,N", [(8, 2048), (3, 257), (5, 1)])
def test_fused_warp_matches_composed(cuda, B, N):
    """pwclo_warp_fwd/_bwd against the op-by-op torch expression of PW/PWCLO_utils.py:31-63 and its autograd:
    forward to 1e-6, gradients wrt q, t and xyz to 1e-5 relative to the largest entry (non-unit quaternions too)"""
    from pwclonet_pylidarslam_b200.pwclonet import PWCLO_utils as U
    g = torch.Generator(device=cuda).manual_seed(B * 1000 + N)
    xyz = (torch.randn(B, 3, N, device=cuda, generator=g) * 10).requires_grad_(True)
    q = (torch.tensor([1.0, 0, 0, 0], device=cuda) + 0.4 * torch.randn(B, 4, device=cuda, generator=g)).reshape(B, 4, 1).requires_grad_(True)
    t = torch.randn(B, 3, 1, device=cuda, generator=g).requires_grad_(True)
    up = torch.randn(B, 3, N, device=cuda, generator=g)
    out = U.FusedWarp.apply(xyz, q, t)
    (out * up).sum().backward()
    got = [out.detach(), xyz.grad.clone(), q.grad.clone(), t.grad.clone()]
    xyz.grad = q.grad = t.grad = None
    ref = U.warp_composed(xyz, q, t)
    (ref * up).sum().backward()
    want = [ref.detach(), xyz.grad, q.grad, t.grad]
    for name, a, r, tol in zip(("out", "dxyz", "dq", "dt"), got, want, (2e-6, 1e-5, 1e-5, 1e-5)):
        assert a.shape == r.shape, name
        err = float((a - r).abs().max())
        assert err <= tol * float(r.abs().max()) + 1e-7, (name, err, float(r.abs().max()))


@pytest.mark.parametrize("B,CI,CO,S,K,need_dx", [(8, 6, 8, 2048, 32, False), (4, 8, 8, 1024, 32, True), (2, 8, 16, 2048, 16, True),
                                                 (3, 3, 8, 4096, 8, True)])
def test_skinny_conv1x1_matches_torch(cuda, B, CI, CO, S, K, need_dx):
    """pwclo_conv1x1_small / pwclo_conv1x1_wgrad against F.conv2d and its autograd (fp32, TF32 off): 1e-5 relative to
    the largest entry on y, dW (a 5e5-term sum) and dx"""
    from pwclonet_pylidarslam_b200.pytorch_utils import SkinnyConv1x1
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device=cuda).manual_seed(CI * 100 + CO)
    x = torch.randn(B, CI, S, K, device=cuda, generator=g).requires_grad_(need_dx)
    w = (torch.randn(CO, CI, 1, 1, device=cuda, generator=g) * 0.4).requires_grad_(True)
    up = torch.randn(B, CO, S, K, device=cuda, generator=g)
    y = SkinnyConv1x1.apply(x, w)
    (y * up).sum().backward()
    got = [y.detach(), w.grad.clone()] + ([x.grad.clone()] if need_dx else [])
    w.grad = None
    if need_dx:
        x.grad = None
    yr = torch.nn.functional.conv2d(x, w)
    (yr * up).sum().backward()
    want = [yr.detach(), w.grad] + ([x.grad] if need_dx else [])
    for name, a, r in zip(("y", "dw", "dx"), got, want):
        err = float((a - r).abs().max())
        assert err <= 1e-5 * float(r.abs().max()) + 1e-7, (name, err, float(r.abs().max()))


def test_ddp_graphed_step_equals_eager_step(cuda):
    """2 GPUs: PWCLONetTrainer.capture with the NCCL all-reduce inside the graph replays the eager data-parallel step
    (same losses, ranks stay identical) and the processes exit cleanly (trainer.close() before destroy_process_group).
    Runs tools/ddp_graph_probe.py under torchrun; skipped on a single-GPU box."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29561", os.path.join(root, "tools", "ddp_graph_probe.py")],
                       capture_output=True, text=True, timeout=240, cwd=root)
    out = r.stdout + r.stderr
    assert r.returncode == 0 and "rank 0: OK" in out and "rank 1: OK" in out, out[-3000:]


@pytest.mark.parametrize("shape", [(2, 16, 300, 32), (3, 64, 100, 16), (2, 64, 77, 8), (1, 5, 33, 6), (2, 3, 10, 1)])
def test_max_over_neighbours_matches_torch(cuda, shape):
    """pwclo_maxpool_lastdim_fwd/_bwd against F.max_pool2d(x, [1, K]) on the same GPU: values, and the gradient routing
    incl. exact ties (ReLU outputs are full of equal zeros: the first maximum must win, as in ATen) and NaN propagation"""
    import torch.nn.functional as F_
    from pwclonet_pylidarslam_b200.pytorch_utils import MaxPoolLastDim
    g = torch.Generator(device=cuda).manual_seed(sum(shape))
    x = torch.relu(torch.randn(shape, device=cuda, generator=g))          # ~half the entries are exactly 0
    x[0, 0, 0] = 0.0                                                      # an all-zero row: K-way tie
    if shape[-1] > 2:
        x[-1, -1, -1, 1] = float("nan")
    up = torch.randn(shape[:-1], device=cuda, generator=g)
    a = x.clone().requires_grad_(True)
    b = x.clone().requires_grad_(True)
    ya = MaxPoolLastDim.apply(a)
    yb = F_.max_pool2d(b, kernel_size=[1, shape[-1]]).squeeze(-1)
    assert torch.equal(torch.nan_to_num(ya, nan=-7.0), torch.nan_to_num(yb, nan=-7.0))
    (ya * up).sum().backward()
    (yb * up).sum().backward()
    assert torch.equal(torch.nan_to_num(a.grad, nan=-7.0), torch.nan_to_num(b.grad, nan=-7.0))


@pytest.mark.parametrize("B,S,K", [(2, 300, 6), (3, 64, 32), (1, 1000, 4)])
def test_cost_geometry_matches_torch_composition(cuda, monkeypatch, B, S, K):
    """pwclo_cost_geometry_fwd/_bwd against the reference's tile / sub / square / sum / sqrt / cat composition
    (PW/costvolume.py:94-105): values bit-exact up to the size-3 summation order (1 ulp on the norm channel), both input
    gradients to 1e-5, incl. the self-neighbour rows where q == p (norm = sqrt(1e-20))"""
    from pwclonet_pylidarslam_b200.pwclonet.costvolume import CostVolume
    g = torch.Generator(device=cuda).manual_seed(B + S + K)
    c0 = torch.randn(B, 3, S, device=cuda, generator=g) * 5
    q0 = c0.unsqueeze(3) + torch.randn(B, 3, S, K, device=cuda, generator=g)
    q0[..., 0] = c0                                    # first neighbour = the point itself
    up = torch.randn(B, 10, S, K, device=cuda, generator=g)
    res = []
    for flag in ("1", "0"):
        monkeypatch.setenv("PWCLO_COST_GEO", flag)
        c, q = c0.clone().requires_grad_(True), q0.clone().requires_grad_(True)
        out = CostVolume._geometry(c, q)
        (out * up).sum().backward()
        res.append((out.detach(), c.grad, q.grad))
    (o1, gc1, gq1), (o0, gc0, gq0) = res
    assert torch.equal(o1[:, :9], o0[:, :9])
    torch.testing.assert_close(o1[:, 9], o0[:, 9], rtol=2e-7, atol=0)
    torch.testing.assert_close(gq1, gq0, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(gc1, gc0, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("shape", [(2, 64, 300, 6), (3, 16, 64, 32), (1, 32, 1000, 4), (2, 8, 33, 8), (1, 4, 50, 16)])
def test_softmax_pool_matches_torch_composition(cuda, monkeypatch, shape):
    """pwclo_softmax_pool_fwd/_bwd (attentive pooling of the cost volume, PW/costvolume.py:139-145, :181-188) against the
    reference's softmax / mul / sum composition on the same GPU in fp32 and against the same expression in fp64: values and
    both input gradients within 1e-5 relative + 5e-6 absolute (a few fp32 ulps of the O(10) terms: grad_w = g p (x - out)
    cancels where x ~ out), and the kernel's distance from the fp64 result is of the size of torch's own fp32 distance;
    incl. rows with large logits (max-subtraction) and a row of equal logits"""
    from pwclonet_pylidarslam_b200 import pytorch_utils as pt
    g = torch.Generator(device=cuda).manual_seed(sum(shape))
    w0 = torch.randn(shape, device=cuda, generator=g) * 3
    w0[0, 0, 0] = 2.5                                  # equal logits: uniform weights
    w0[0, 0, 1] += 80.0                                # large logits: exp() needs the max subtracted
    x0 = torch.relu(torch.randn(shape, device=cuda, generator=g))
    up = torch.randn(shape[:-1], device=cuda, generator=g)
    res = []
    for flag in ("1", "0"):
        monkeypatch.setenv("PWCLO_SOFTMAX_POOL", flag)
        w, x = w0.clone().requires_grad_(True), x0.clone().requires_grad_(True)
        out = pt.softmax_pool(w, x)
        (out * up).sum().backward()
        res.append((out.detach(), w.grad, x.grad))
    w, x = w0.double().requires_grad_(True), x0.double().requires_grad_(True)
    out = torch.sum(torch.softmax(w, dim=3) * x, dim=3)
    (out * up.double()).sum().backward()
    exact = (out.detach(), w.grad, x.grad)
    for ours, torch32, want, name in zip(res[0], res[1], exact, ("out", "grad_w", "grad_x")):
        e_ours = (ours.double() - want).abs().max().item()
        e_torch = (torch32.double() - want).abs().max().item()
        print(f"softmax_pool {shape} {name}: |ours - fp64| {e_ours:.2e}, |torch fp32 - fp64| {e_torch:.2e}, "
              f"|ours - torch fp32| {(ours - torch32).abs().max().item():.2e}")
        torch.testing.assert_close(ours, torch32, rtol=1e-5, atol=5e-6, msg=lambda m: f"{name}: {m}")
        assert e_ours <= 3 * e_torch + 1e-6, (name, e_ours, e_torch)
    with torch.no_grad():                              # without autograd: the reference's expression itself
        monkeypatch.setenv("PWCLO_SOFTMAX_POOL", "1")
        assert torch.equal(pt.softmax_pool(w0, x0), torch.sum(torch.softmax(w0, dim=3) * x0, dim=3))
