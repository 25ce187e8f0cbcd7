"""Pose post-processing rows (N2 / N4) on the GPU against the oracle: quat2mat rotation entries bit-exact
(fp32 op for op), inverse and chained absolute poses to 1e-12 relative (adjugate inverse / prefix product
instead of LAPACK LU and a sequential chain), and the odometry adapter end to end."""
import numpy as np
import pytest
import torch

from oracle import pose_port as P
from pwclonet_pylidarslam_b200 import odometry as O, synthetic as syn

pytestmark = pytest.mark.gpu


def _params(B, seed, scale=0.05):
    rng = np.random.default_rng(seed)
    t = rng.normal(0, 1, (B, 4, 3))
    q = np.array([1.0, 0, 0, 0]) + scale * rng.standard_normal((B, 4, 4))
    return np.concatenate([t, q], axis=-1).astype(np.float32)


def test_pose_to_matrix(cuda):
    pp = _params(300, 0, scale=0.6)
    pp[7, 0, 3:] = 0                                  # degenerate quaternion -> identity rotation
    T = O.pose_params_to_matrices(torch.from_numpy(pp).to(cuda), invert=False).cpu().numpy()
    Ti = O.pose_params_to_matrices(torch.from_numpy(pp).to(cuda), invert=True).cpu().numpy()
    for b in range(300):
        want = P.relative_pose(pp[b, 0], invert=False)
        np.testing.assert_array_equal(T[b], want, err_msg=f"frame {b}")          # fp32 quat2mat, bit-exact
        np.testing.assert_allclose(Ti[b], np.linalg.inv(want), rtol=0, atol=1e-12 * max(1.0, np.abs(want).max()))
    flat = O.pose_params_to_matrices(torch.from_numpy(np.ascontiguousarray(pp[:, 0])).to(cuda), invert=False).cpu().numpy()
    np.testing.assert_array_equal(flat, T)


@pytest.mark.parametrize("F", [1, 31, 1024, 4541])
def test_accumulate_poses(cuda, F):
    pp = _params(F, F)
    pp[:, 0, :3] = np.random.default_rng(5).normal([0, 0, 1.0], 0.1, (F, 3))       # ~1 m per frame: km-long tracks
    rel = np.stack([P.relative_pose(pp[i, 0]) for i in range(F)])
    want = P.convert_to_absolute(rel)
    got = O.convert_to_absolute(torch.from_numpy(rel).to(cuda)).cpu().numpy()
    scale = max(1.0, np.abs(want).max())
    assert np.abs(got - want).max() <= 1e-11 * scale
    first = P.relative_pose(_params(1, 99)[0, 0])
    got = O.convert_to_absolute(torch.from_numpy(rel).to(cuda), first, dict_semantics=True).cpu().numpy()
    assert np.abs(got - P.convert_to_absolute(rel, first)).max() <= 1e-11 * scale
    # default = the reference's ndarray branch (the input IS an array): inv(rel_f ... rel_0 @ first)
    got = O.convert_to_absolute(torch.from_numpy(rel).to(cuda), first).cpu().numpy()
    assert np.abs(got - P.convert_to_absolute_array(rel, first)).max() <= 1e-10 * scale


def test_odometry_adapter(cuda):
    odo = O.PWCLONetOdometry({"num_points": 8192})
    odo.init()
    frames = [syn.make_pair(800 + i, 8192) for i in range(2)]
    clouds = [frames[0]["pc2"], frames[0]["pc1"], frames[1]["pc1"]]
    for c in clouds:
        d = {"odometry_pc": c}
        odo.process_next_frame(d)
        assert d["odometry_pose"].shape == (4, 4)
    rel = odo.get_relative_poses()
    assert rel.shape == (3, 4, 4) and (rel[0] == np.eye(4)).all() and np.isfinite(rel).all()
    # same numbers as running the module and the reference post-processing by hand
    with torch.no_grad():
        a = torch.from_numpy(clouds[1]).to(cuda).unsqueeze(0)
        b = torch.from_numpy(clouds[0]).to(cuda).unsqueeze(0)
        pose, _ = odo.prediction_module([a, b])
    want = P.relative_pose(pose[0, 0].cpu().numpy())
    np.testing.assert_allclose(rel[1], want, atol=1e-12)
    absolute = odo.get_absolute_poses()
    np.testing.assert_allclose(absolute, P.convert_to_absolute(rel), atol=1e-11)
    # batched offline form gives the same relative poses
    odo2 = O.PWCLONetOdometry({"num_points": 8192}, prediction_module=odo.prediction_module)
    odo2.init()
    rel2 = odo2.process_pairs(torch.from_numpy(np.stack(clouds)).to(cuda)).cpu().numpy()
    np.testing.assert_allclose(rel2, rel, atol=1e-6)
    assert len(odo.elapsed) == 3 and odo.get_elapsed() > 0
