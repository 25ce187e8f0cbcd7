"""oracle/pose_port.py -- TEST INFRASTRUCTURE ONLY (checker for the pose post-processing rows N2 / N4).

The reference's host code restated as plain functions: `quat2mat` (train.py:762-796), the relative pose
of one prediction (train.py:875-886: [R|t] then np.linalg.inv) and `convert_to_absolute`, dict and
ndarray branches (slam/common/kitti360_utils.py:406-432).  Pinned in tests/test_pose_cpu.py against the reference's own
`convert_to_absolute` (imported from /root/reference where mounted) and against rotation-matrix identities.
"""
import numpy as np


def quat2mat(q):
    w, x, y, z = q
    Nq = w * w + x * x + y * y + z * z
    if Nq < 1e-8:
        return np.eye(3)
    s = 2.0 / Nq
    X, Y, Z = x * s, y * s, z * s
    wX, wY, wZ = w * X, w * Y, w * Z
    xX, xY, xZ = x * X, x * Y, x * Z
    yY, yZ, zZ = y * Y, y * Z, z * Z
    return np.array([[1.0 - (yY + zZ), xY - wZ, xZ + wY], [xY + wZ, 1.0 - (xX + zZ), yZ - wX], [xZ - wY, yZ + wX, 1.0 - (xX + yY)]])


def relative_pose(pred_params, invert=True):
    """pred_params float32[7] = (t, q) -> float64 4x4 (train.py:875-886)"""
    q = pred_params[3:].reshape(4)
    t = pred_params[:3].reshape((3, 1))
    R = quat2mat(q)
    T = np.concatenate([np.concatenate([R, t], axis=-1), np.array([[0.0, 0.0, 0.0, 1.0]])], axis=0)
    return np.linalg.inv(T) if invert else T


def convert_to_absolute(relative_poses, first_transformation=None):
    """dict branch, kitti360_utils.py:422-427"""
    prev = np.eye(4) if first_transformation is None else first_transformation
    out = []
    for rel in relative_poses:
        prev = np.linalg.inv(rel @ np.linalg.inv(prev))
        out.append(prev)
    return np.stack(out)


def convert_to_absolute_array(relative_poses, first_transformation=None):
    """ndarray branch, kitti360_utils.py:412-420: accumulate rel_i @ prev, invert everything at the end"""
    prev = np.eye(4) if first_transformation is None else first_transformation
    out = []
    for rel in relative_poses:
        prev = rel @ prev
        out.append(prev)
    return np.linalg.inv(np.stack(out))
