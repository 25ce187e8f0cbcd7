"""oracle/make_golden_train.py -- TEST INFRASTRUCTURE ONLY.  Runs the UNMODIFIED reference loss module
(slam/training/loss_modules.py:329-544, via oracle/ref_shim.load_reference_loss) on CPU and writes
tests/golden/loss_kat.npz: inputs, loss, the reference's log values and autograd gradients.
Needs /root/reference; run once here, the fixture travels.   python -m oracle.make_golden_train"""
import os

import numpy as np
import torch

from oracle import ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "loss_kat.npz")
ORDER = (["loss"] + [f"loss_l{i}" for i in range(1, 5)] + [f"loss_rot_l{i}" for i in range(1, 5)]
         + [f"loss_trans_l{i}" for i in range(1, 5)])


def case(rng, B, kind):
    gt = rng.standard_normal((B, 7)).astype(np.float32)
    gt[:, 3:] /= np.linalg.norm(gt[:, 3:], axis=-1, keepdims=True)
    if kind == "random":
        pred = rng.standard_normal((B, 4, 7)).astype(np.float32)
    elif kind == "near":          # a trained network: small errors, exercises the 1e-10 guards
        pred = (gt[:, None, :] + 1e-3 * rng.standard_normal((B, 4, 7))).astype(np.float32)
    else:                         # exact: pred == gt on some entries (sqrt(0 + 1e-10) branches)
        pred = np.repeat(gt[:, None, :], 4, axis=1).copy()
        pred[:, 1:, :] += (0.1 * rng.standard_normal((B, 3, 7))).astype(np.float32)
    return pred, gt


def main():
    rng = np.random.default_rng(20240)
    out = {}
    cases = [(1, "random", True, (0.0, -2.5)), (5, "random", True, (0.0, -2.5)), (8, "near", True, (0.3, -1.7)),
             (64, "random", True, (-0.5, -3.0)), (7, "exact", True, (0.0, -2.5)), (6, "random", False, (1.0, 1.0)),
             (300, "near", False, (0.5, 2.0))]
    for i, (B, kind, with_exp, s) in enumerate(cases):
        pred, gt = case(rng, B, kind)
        mod = ref_shim.load_reference_loss(with_exp_weights=with_exp, init_weights=s, loss_weights=s)
        p = torch.tensor(pred, requires_grad=True)
        loss, log = mod(p, torch.tensor(gt))
        loss.backward()
        out[f"c{i}_pred"], out[f"c{i}_gt"], out[f"c{i}_s"] = pred, gt, np.asarray(s, np.float32)
        out[f"c{i}_with_exp"] = np.asarray(int(with_exp))
        out[f"c{i}_terms"] = np.asarray([float(log[k].detach()) for k in ORDER], np.float32)
        out[f"c{i}_grad_pred"] = p.grad.numpy().copy()
        if with_exp:
            out[f"c{i}_grad_s"] = mod.exp_weighting.s_param.grad.numpy().copy()
    out["n_cases"] = np.asarray(len(cases))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
