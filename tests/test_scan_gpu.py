"""Input-pipeline row (N3) on the GPU: pwclo_prepare_scans against the numpy oracle, bit-exact (float64
transform in np.matmul's accumulation order, integer key selection)."""
import numpy as np
import pytest
import torch

from oracle import scan_port as S
from pwclonet_pylidarslam_b200 import scan_input, synthetic as syn

pytestmark = pytest.mark.gpu


def _run(cuda, scans, npoints, seed, post=None, Tr=None):
    buf, off, mx = scan_input.pack_scans(scans)
    Tr = syn.KITTI_TR if Tr is None else Tr
    out, idx, surv = scan_input.prepare_scans(buf.to(cuda), off.to(cuda), torch.from_numpy(np.ascontiguousarray(Tr)).to(cuda),
                                              npoints, seed, post=None if post is None else torch.from_numpy(post).to(cuda),
                                              max_points=mx, return_index=True)
    torch.cuda.synchronize()
    return out.cpu().numpy(), idx.cpu().numpy(), surv.cpu().numpy()


def _check(scans, npoints, seed, got, post=None, Tr=None):
    out, idx, surv = got
    for s, raw in enumerate(scans):
        T = syn.KITTI_TR if Tr is None else (Tr[s] if Tr.ndim == 3 else Tr)
        want, sel, M = S.prepare_scan(raw, T, s, seed, npoints, None if post is None else post[s])
        assert surv[s] == M, (s, surv[s], M)
        np.testing.assert_array_equal(idx[s], sel, err_msg=f"scan {s}")
        np.testing.assert_array_equal(out[s], want, err_msg=f"scan {s}")


def test_full_scans_bit_exact(cuda):
    scans = [syn.make_raw_scan(11), syn.make_raw_scan(12), syn.make_raw_scan(13)[:70001], syn.make_raw_scan(11)[:9000]]
    got = _run(cuda, scans, 8192, seed=0x1234567890ABCDEF)
    _check(scans, 8192, 0x1234567890ABCDEF, got)
    # no duplicates where there were enough survivors; every chosen point passes the crop
    for s in range(3):
        assert len(set(got[1][s].tolist())) == 8192
        assert (got[0][s][:, 1] <= 1.1).all() and (np.abs(got[0][s][:, [0, 2]]) < 30).all()


def test_short_empty_and_tiny_scans(cuda):
    rng = np.random.default_rng(0)
    far = rng.uniform(-1, 1, size=(300, 4)).astype(np.float32)
    far[:, 0] += 100.0
    some = far.copy()
    some[:40, 0] -= 95.0
    scans = [some, far, syn.make_raw_scan(14)[:5], syn.make_raw_scan(14)[:8192 * 2 + 3]]
    got = _run(cuda, scans, 8192, seed=7)
    _check(scans, 8192, 7, got)
    assert got[2][1] == 0 and got[2][0] == 40


def test_augmentation_transform_and_16384_points(cuda):
    scans = [syn.make_raw_scan(15), syn.make_raw_scan(16)]
    rng = np.random.default_rng(1)
    post = np.tile(np.eye(4)[:3], (2, 1, 1))
    a = 0.03
    post[1, :3, :3] = [[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]
    post[1, :3, 3] = rng.normal(0, 0.3, 3)
    post = np.ascontiguousarray(post)
    got = _run(cuda, scans, 16384, seed=99, post=post)
    _check(scans, 16384, 99, got, post=post)
    # per-scan calibration matrices
    Tr = np.ascontiguousarray(np.stack([syn.KITTI_TR, syn.KITTI_TR * 1.0]))
    Tr[1, :, 3] += 0.01
    got = _run(cuda, scans, 4096, seed=5, Tr=Tr)
    _check(scans, 4096, 5, got, Tr=Tr)


def test_kitti360_crop(cuda):
    """KITTI-360 flavour: velodyne frame, identity Tr, ground below the wheel axis, 16384 points kept"""
    scans = [syn.make_raw_scan(31), syn.make_raw_scan(32)[:60000]]
    thr = np.float32(-(1.73 - 0.3))
    scans[0][:64, 2] = thr
    scans[0][64:128, 2] = np.nextafter(thr, np.float32(-10))
    crop = scan_input.kitti360_crop(30.0)
    buf, off, mx = scan_input.pack_scans(scans)
    out, idx, surv = scan_input.prepare_scans(buf.to(cuda), off.to(cuda), torch.from_numpy(scan_input.IDENTITY_TR.copy()).to(cuda),
                                              16384, 77, return_index=True, crop=crop)
    out, idx, surv = out.cpu().numpy(), idx.cpu().numpy(), surv.cpu().numpy()
    for s, raw in enumerate(scans):
        want, sel, M = S.prepare_scan(raw, scan_input.IDENTITY_TR, s, 77, 16384, crop=crop)
        assert surv[s] == M == S.reference_mask_kitti360(raw[:, :3]).sum()
        np.testing.assert_array_equal(idx[s], sel)
        np.testing.assert_array_equal(out[s], want)


def test_pipeline_feeds_the_network(cuda):
    """raw scans -> prepare_scans -> PWCLONet fused forward, no host round trip in between"""
    from pwclonet_pylidarslam_b200.pwclonet import PWCLONet
    a, b = syn.make_raw_scan(21), syn.make_raw_scan(22)
    buf, off, mx = scan_input.pack_pairs([a], [b])
    clouds = scan_input.prepare_scans(buf.to(cuda), off.to(cuda), torch.from_numpy(syn.KITTI_TR.copy()).to(cuda), 8192, 3,
                                      max_points=mx)
    net = PWCLONet({"device": "cuda:0"}).to(cuda).eval()
    with torch.no_grad():
        pose, _ = net(clouds[0:1].permute(0, 2, 1).contiguous(), None, clouds[1:2].permute(0, 2, 1).contiguous(), None)
    assert pose.shape == (1, 4, 7) and torch.isfinite(pose).all()


def test_refuses_cpu_tensors():
    with pytest.raises(RuntimeError, match="no CPU path"):
        scan_input.prepare_scans(torch.zeros(8, 4), torch.tensor([0, 8]), torch.zeros(3, 4, dtype=torch.float64), 4, 0)
