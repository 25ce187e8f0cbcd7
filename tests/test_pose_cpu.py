"""CPU suite for the pose post-processing rows (N2 / N4): the oracle against the reference's own
convert_to_absolute and against rotation identities."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import pose_port as P

REF = "/root/reference/slam/common/kitti360_utils.py"


def test_quat2mat_is_a_rotation_and_matches_known_answers():
    np.testing.assert_allclose(P.quat2mat(np.array([1, 0, 0, 0], np.float32)), np.eye(3))
    # 90 degrees about z
    c = np.float32(np.sqrt(0.5))
    np.testing.assert_allclose(P.quat2mat(np.array([c, 0, 0, c], np.float32)), [[0, -1, 0], [1, 0, 0], [0, 0, 1]], atol=1e-6)
    rng = np.random.default_rng(0)
    for _ in range(20):
        q = rng.standard_normal(4).astype(np.float32)          # non-unit quaternions are allowed
        R = P.quat2mat(q)
        np.testing.assert_allclose(R @ R.T, np.eye(3), atol=5e-6)
        assert abs(np.linalg.det(R) - 1) < 5e-6
    assert (P.quat2mat(np.zeros(4, np.float32)) == np.eye(3)).all()


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not mounted")
def test_convert_to_absolute_equals_reference():
    src = open(REF).read()
    # the function needs only numpy: execute the class that holds it without the module's heavy imports
    start = src.index("class KITTI360_TRANSFORMATIONS")
    end = src.index("class KITTI360_TOOLS")
    ns = {"np": np}
    exec(compile("import numpy as np\n" + src[start:end], REF, "exec"), ns)
    ref = ns["KITTI360_TRANSFORMATIONS"].convert_to_absolute
    rng = np.random.default_rng(1)
    rel = {i: P.relative_pose(np.concatenate([rng.normal(0, 1, 3), [1, 0, 0, 0] + 0.05 * rng.standard_normal(4)]).astype(np.float32))
           for i in range(50)}
    want = ref(rel)
    got = P.convert_to_absolute([rel[i] for i in range(50)])
    for i in range(50):
        np.testing.assert_array_equal(got[i], want[i])
    # ndarray branch, with and without a non-identity first transformation (the two branches differ there)
    arr = np.stack([rel[i] for i in range(50)])
    first = P.relative_pose(np.array([0.3, -0.2, 1.5, 0.98, 0.05, -0.1, 0.02], np.float32))
    np.testing.assert_array_equal(P.convert_to_absolute_array(arr), ref(arr))
    np.testing.assert_array_equal(P.convert_to_absolute_array(arr, first), ref(arr, first))
    wd = ref(rel, first)
    gd = P.convert_to_absolute([rel[i] for i in range(50)], first)
    for i in range(50):
        np.testing.assert_array_equal(gd[i], wd[i])
    assert np.abs(gd - P.convert_to_absolute_array(arr, first)).max() > 1e-3      # genuinely different semantics


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not mounted")
def test_save_poses_equals_reference_text(tmp_path):
    """the KITTI pose text file, byte for byte, against the reference's KITTI360_IO.save_poses (pandas to_csv)"""
    import pandas as pd
    from pwclonet_pylidarslam_b200 import odometry as O
    src = open(REF).read()
    start = src.index("class KITTI360_IO")
    end = src.index("class KITTI360_TRANSFORMATIONS")
    body = src[start:end]
    # only save_poses is needed: cut the class down to that method
    a = body.index("    def save_poses")
    b = body.index("    def loadWindowNpy")
    ns = {"np": np, "pd": pd}
    exec(compile("class KITTI360_IO:\n" + body[a:b], REF, "exec"), ns)
    ref = ns["KITTI360_IO"].save_poses
    rng = np.random.default_rng(2)
    rel = [P.relative_pose(np.concatenate([rng.normal(0, 1, 3), [1, 0, 0, 0] + 0.05 * rng.standard_normal(4)]).astype(np.float32))
           for _ in range(40)]
    absolute = P.convert_to_absolute(rel)
    absolute[0] = np.eye(4)                                   # exact 1.0 / 0.0 entries
    absolute[1][0, 3] = 1e-17
    absolute[2][1, 3] = 123456789.125
    d = {3 * i + 7: absolute[i] for i in range(40)}           # dict form with sparse frame ids
    ref(d, str(tmp_path / "ref_dict.txt"))
    O.save_poses(d, str(tmp_path / "our_dict.txt"))
    assert open(tmp_path / "ref_dict.txt").read() == open(tmp_path / "our_dict.txt").read()
    ref(absolute, str(tmp_path / "ref_arr.txt"))
    O.save_poses(absolute, str(tmp_path / "our_arr.txt"))
    assert open(tmp_path / "ref_arr.txt").read() == open(tmp_path / "our_arr.txt").read()
