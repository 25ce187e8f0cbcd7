"""oracle/train_port.py -- TEST INFRASTRUCTURE ONLY (checker for the training-side rows; never imported by
the product package).

CPU restatement, in torch fp32 on the host, of
  * `_PWCLONetLossModule.forward` + `ExponentialWeights.forward`
    (slam/training/loss_modules.py:424-544, :171-196) -- `pose_loss`, gradient by torch autograd;
  * one `torch.optim.Adam` step as the reference configures it (slam/training/trainer.py:309-323:
    betas (0.9, 0.999), weight_decay added to the gradient) -- `adam_step`, written out in numpy.

Pinned: `pose_loss` equals the UNMODIFIED reference loss module (imported through
oracle/ref_shim.load_reference_loss) bit for bit on CPU -- tests/test_training_cpu.py, live where
/root/reference is mounted and through tests/golden/loss_kat.npz (oracle/make_golden_train.py)
elsewhere; `adam_step` is checked against torch.optim.Adam itself in the same test file.
"""
import numpy as np
import torch


def _norm(x):                                         # loss_modules.py:388-391
    return x / (torch.sqrt(torch.sum(x * x, dim=-1, keepdim=True) + 1e-10) + 1e-10)


def _l2_norm(x, g):                                   # :370-373
    return torch.mean(torch.sqrt(torch.sum((x - g) * (x - g), dim=-1, keepdim=True) + 1e-10))


def _trans(x, g):                                     # :382-384
    return torch.mean(torch.sqrt((x - g) * (x - g) + 1e-10))


def pose_loss(pred, gt, s, with_exp=True, need_grad=True):
    """pred[B,4,7], gt[B,7], s[2] (numpy fp32) -> dict(loss, terms[16] in the layout of
    pwclo_pose_loss, grad_pred, grad_s)."""
    p = torch.tensor(np.asarray(pred, np.float32), requires_grad=need_grad)
    sp = torch.tensor(np.asarray(s, np.float32), requires_grad=need_grad)
    g = torch.tensor(np.asarray(gt, np.float32))
    rot, tr, lvl = [], [], []
    for l in range(4):                                # :446-479, row 0 = level 1 (finest)
        r = _l2_norm(_norm(p[:, l, 3:]), g[:, 3:])
        t = _trans(p[:, l, :3], g[:, :3])
        rot.append(r)
        tr.append(t)
        if with_exp:                                  # ExponentialWeights([trans, rot]) :171-196
            L = 0.0
            for x, si in ((t, sp[0]), (r, sp[1])):
                L = L + (x * torch.exp(-si) + si)
        else:                                         # :519-522
            L = t * sp[0] + r * sp[1]
        lvl.append(L)
    loss = 1.6 * lvl[3] + 0.8 * lvl[2] + 0.4 * lvl[1] + 0.2 * lvl[0]      # :531
    out = {"loss": float(loss.detach())}
    terms = np.zeros(16, np.float32)
    terms[0] = loss.detach().numpy()
    for l in range(4):
        terms[1 + l] = lvl[l].detach().numpy()
        terms[5 + l] = rot[l].detach().numpy()
        terms[9 + l] = tr[l].detach().numpy()
    terms[13:15] = np.asarray(s, np.float32)
    terms[15] = p.shape[0]
    out["terms"] = terms
    if need_grad:
        loss.backward()
        out["grad_pred"] = p.grad.numpy().copy()
        out["grad_s"] = sp.grad.numpy().copy()
    return out


def adam_step(p, g, m, v, step, lr, b1=0.9, b2=0.999, eps=1e-8, wd=0.0, grad_scale=1.0):
    """one torch.optim.Adam update, fp32 state, bias corrections in python floats (torch/optim/adam.py
    _single_tensor_adam).  Returns new (p, m, v)."""
    f = np.float32
    g = g.astype(f) * f(grad_scale)
    g = g + f(wd) * p
    m = m + (g - m) * f(1.0 - b1)
    v = v * f(b2) + (g * g) * f(1.0 - b2)
    bc1, bc2 = 1.0 - b1 ** step, 1.0 - b2 ** step
    denom = np.sqrt(v) / f(np.sqrt(bc2)) + f(eps)
    p = p - f(lr / bc1) * (m / denom)
    return p.astype(f), m.astype(f), v.astype(f)
