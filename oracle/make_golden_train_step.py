"""oracle/make_golden_train_step.py -- TEST INFRASTRUCTURE ONLY.  One TRAINING step of the UNMODIFIED reference on CPU:
reference PWCLONet in train() (batch statistics in every BatchNorm, running-statistics update) through oracle/ref_shim
(CPU `_ext` = oracle/pointnet2_cpu.c, incl. the group_points_grad scatter-add of EXT/src/group_points_gpu.cu:43-64), the
UNMODIFIED loss module (slam/training/loss_modules.py:329-544), `loss.backward()` by torch autograd.
The two dropouts of every pose head (PW/pose_calculator.py:63,65) are patched to the identity: their masks come from
torch's generator and cannot be replayed by another implementation.
Writes tests/golden/train_step_b2_n2048.npz: loss, pose, gradients of a spread of named tensors (every kind of layer on
the path), updated BN running statistics.  Needs /root/reference; run once here, the fixture travels.
    python -m oracle.make_golden_train_step"""
import hashlib
import os

import numpy as np
import torch
import torch.nn.functional as F

from oracle import ref_shim
from pwclonet_pylidarslam_b200 import synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "train_step_b2_n2048.npz")
FIRST_PAIR, PAIRS, POINTS, WSEED = 930, 2, 2048, 4
GRADS = ["psa_1.mlp_module.layer0.conv.weight", "psa_1.mlp_module.layer2.bn.bn.weight", "psa_2.mlp_module.layer1.conv.weight",
         "psa_3.mlp_module.layer0.conv.weight", "psa_4.mlp_module.layer2.bn.bn.bias", "cost_volume.mlp_convs.layer0.conv.weight",
         "cost_volume.mlp_conv_xyz_1.layer0.conv.weight", "cost_volume.mlp3_convs.layer1.conv.weight",
         "flow_feature_encoding.mlp_module.layer1.conv.weight", "l4_flow_predictor.mlp_convs.layer0.conv.weight",
         "pose_calculator_4.conv1d_q.conv.weight", "pose_calculator_4.conv1d_q_t.conv.bias",
         "pose_warp_refinement_3.setupconv_features.mlp.layer0.conv.weight",
         "pose_warp_refinement_3.cost_volume.mlp3_convs.layer0.conv.weight",
         "pose_warp_refinement_2.setupconv_mask.post_mlp.layer0.conv.weight",
         "pose_warp_refinement_2.cost_volume.mlp2_convs.layer1.bn.bn.weight",
         "pose_warp_refinement_2.flow_predictor_mask.mlp_convs.layer1.conv.weight",
         "pose_warp_refinement_1.cost_volume.mlp_convs.layer2.conv.weight",
         "pose_warp_refinement_1.flow_predictor_features.mlp_convs.layer0.conv.weight",
         "pose_warp_refinement_1.pose_calculator.conv1d_t.conv.weight"]
STATS = ["psa_1.mlp_module.layer0.bn.bn", "psa_3.mlp_module.layer2.bn.bn", "cost_volume.mlp2_convs.layer0.bn.bn",
         "pose_warp_refinement_2.setupconv_features.mlp.layer1.bn.bn", "pose_warp_refinement_1.cost_volume.mlp3_convs.layer1.bn.bn"]


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    F.dropout = lambda x, p=0.5, training=True, inplace=False: x
    torch.set_num_threads(os.cpu_count() or 1)
    net = ref_shim.load_reference()
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    w = syn.make_state_dict(shapes, seed=WSEED)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    net.train()
    loss_mod = ref_shim.load_reference_loss()
    x1, x2, gt = syn.make_batch(FIRST_PAIR, PAIRS, POINTS)
    pose, _ = net(torch.from_numpy(x1), None, torch.from_numpy(x2), None)
    loss, log = loss_mod(pose, torch.from_numpy(gt))
    loss.backward()
    params = dict(net.named_parameters())
    sd = net.state_dict()
    out = {"meta": np.asarray([FIRST_PAIR, PAIRS, POINTS, WSEED]), "input_sha": np.asarray(sha(x1, x2, gt)),
           "weights_sha": np.asarray(sha(*[w[k] for k in sorted(w)])), "loss": np.asarray(float(loss), np.float32),
           "pose": pose.detach().numpy(), "grad_s": loss_mod.exp_weighting.s_param.grad.numpy().copy(),
           "grad_names": np.asarray(GRADS), "stat_names": np.asarray(STATS),
           "grad_norm_all": np.asarray([float(p.grad.double().norm()) for p in params.values()], np.float64)}
    for i, n in enumerate(GRADS):
        out[f"grad_{i}"] = params[n].grad.numpy().copy()
    for i, n in enumerate(STATS):
        out[f"mean_{i}"] = sd[n + ".running_mean"].numpy().copy()
        out[f"var_{i}"] = sd[n + ".running_var"].numpy().copy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes; loss", float(loss))


if __name__ == "__main__":
    main()
