"""CPU suite for the input-pipeline row (N3): the oracle's generator against Random123's known answers, its
transform + crop against the literal reference expression, its selection against its definition."""
import numpy as np

from oracle import scan_port as S
from pwclonet_pylidarslam_b200 import synthetic as syn


def test_philox_known_answers():
    """Random123 kat_vectors, philox4x32-10"""
    u = np.uint32
    for ctr, key, want in (((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
                           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
                           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
                            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))):
        got = S.philox4x32_10(*[u(c) for c in ctr], *key)
        assert tuple(int(g) for g in got) == want


def test_transform_and_mask_equal_reference_expression():
    raw = syn.make_raw_scan(5)
    assert raw.dtype == np.float32 and raw.shape[1] == 4 and raw.shape[0] > 50000
    Tr4 = np.vstack([syn.KITTI_TR, [0, 0, 0, 1.0]])
    ref_pts, ref_mask = S.reference_transform_and_mask(raw, Tr4)
    P = S.affine(syn.KITTI_TR, raw[:, :3])
    # bit for bit in float64 (np.matmul = multiply + fused multiply-adds in storage order); a BLAS with another
    # accumulation order would still agree to an ulp of float64
    assert np.abs(P - ref_pts).max() <= 4e-14
    np.testing.assert_array_equal(P, ref_pts)
    np.testing.assert_array_equal(S.keep_mask(P), ref_mask)
    assert 8192 < ref_mask.sum() < raw.shape[0]


def test_selection_is_a_subset_without_replacement_in_key_order():
    raw = syn.make_raw_scan(6)
    pts, sel, M = S.prepare_scan(raw, syn.KITTI_TR, scan=3, seed=1234, npoints=8192)
    mask = S.keep_mask(S.affine(syn.KITTI_TR, raw[:, :3]))
    assert M == mask.sum() and len(set(sel.tolist())) == 8192 and mask[sel].all()
    k = S.point_keys(raw.shape[0], 3, 1234)
    comp = (k[sel].astype(np.uint64) << np.uint64(32)) | sel.astype(np.uint64)
    assert (np.diff(comp.astype(np.float64)) > 0).all()
    # nothing outside the sample has a smaller key than the largest key inside
    rest = np.setdiff1d(np.where(mask)[0], sel)
    assert k[rest].min() >= k[sel].max()
    np.testing.assert_array_equal(pts, S.affine(syn.KITTI_TR, raw[sel, :3]).astype(np.float32))
    # a different seed or scan id gives a different sample
    _, sel2, _ = S.prepare_scan(raw, syn.KITTI_TR, scan=4, seed=1234, npoints=8192)
    assert len(set(sel.tolist()) & set(sel2.tolist())) < 4096


def test_selection_is_uniform_over_survivors():
    """every survivor equally likely: chi-square of inclusion counts over many seeds (definition of
    np.random.choice(replace=False), kitti_odometry_dataset.py:163)"""
    mask = np.zeros(400, bool)
    mask[::2] = True
    counts = np.zeros(400)
    trials = 600
    for seed in range(trials):
        sel, _ = S.select(mask, 0, seed, 50)
        counts[sel] += 1
    assert counts[~mask].sum() == 0
    exp = trials * 50 / 200
    chi2 = ((counts[mask] - exp) ** 2 / (exp * (1 - 50 / 200))).sum()
    assert 130 < chi2 < 280          # 199 degrees of freedom: mean 199, sd 20


def test_short_and_empty_scans():
    rng = np.random.default_rng(0)
    raw = rng.uniform(-1, 1, size=(300, 4)).astype(np.float32)
    raw[:, 0] += 100.0                            # far ahead (z_cam = x_velo): everything is cropped
    raw[:40, 0] -= 95.0                           # 40 survivors
    pts, sel, M = S.prepare_scan(raw, syn.KITTI_TR, 0, 9, 128)
    assert M == 40 and sel[:40].tolist() == list(range(40)) and set(sel[40:].tolist()) <= set(range(40))
    pts, sel, M = S.prepare_scan(raw[40:], syn.KITTI_TR, 0, 9, 128)
    assert M == 0 and sel.min() >= 0 and sel.max() < 260 and len(set(sel.tolist())) > 60


def test_kitti360_crop_equals_reference_expression():
    """velodyne-frame crop of kitti_360_dataset_2.py:113-123 (float32 comparisons): float32-rounded thresholds make
    the float64 comparisons of oracle and kernel decide identically, also for coordinates exactly at a threshold"""
    from pwclonet_pylidarslam_b200 import scan_input
    raw = syn.make_raw_scan(8)
    crop = scan_input.kitti360_crop(30.0)
    thr = np.float32(-(1.73 - 0.3))
    raw[:50, 2] = thr                                   # exactly on the ground threshold
    raw[50:100, 2] = np.nextafter(thr, np.float32(-10))
    raw[100:150, 2] = np.nextafter(thr, np.float32(10))
    raw[150:170, 0] = np.float32(30.0)
    P = S.affine(scan_input.IDENTITY_TR, raw[:, :3])
    np.testing.assert_array_equal(P, raw[:, :3].astype(np.float64))          # identity Tr is exact
    np.testing.assert_array_equal(S.keep_mask(P, crop), S.reference_mask_kitti360(raw[:, :3], 30.0))
    pts, sel, M = S.prepare_scan(raw, scan_input.IDENTITY_TR, 0, 3, 16384, crop=crop)
    assert M == S.reference_mask_kitti360(raw[:, :3]).sum() and pts.shape == (16384, 3)
    np.testing.assert_array_equal(pts, raw[sel, :3])


def test_pack_scans_and_pairs_host_side():
    """host packing: one buffer + int64 offsets; a frame pair is cut to the shorter scan first (reference :378-384)"""
    from pwclonet_pylidarslam_b200 import scan_input
    rng = np.random.default_rng(0)
    a = rng.standard_normal((1000, 4)).astype(np.float32)
    b = rng.standard_normal((700, 4)).astype(np.float32)
    c = rng.standard_normal((5, 4)).astype(np.float32)
    buf, off, mx = scan_input.pack_scans([a, b, c], pin=False)
    assert buf.shape == (1705, 4) and off.tolist() == [0, 1000, 1700, 1705] and mx == 1000
    np.testing.assert_array_equal(buf[1000:1700].numpy(), b)
    buf, off, mx = scan_input.pack_pairs([a, c], [b, a], pin=False)
    assert off.tolist() == [0, 700, 1400, 1405, 1410] and mx == 700
    np.testing.assert_array_equal(buf[:700].numpy(), a[:700])
    np.testing.assert_array_equal(buf[1405:1410].numpy(), a[:5])
    # crop helpers: KITTI-360 thresholds are float32-rounded, the KITTI odometry ones are the reference literals
    g = scan_input.kitti360_crop(30.0)
    assert g[:2] == (2, -1) and g[2] == float(np.float32(-1.43)) and g[3:5] == (0, 1) and g[5] == 30.0
    assert scan_input.KITTI_ODOMETRY_CROP == (1, 1, 1.1, 0, 2, 30.0)


def test_kitti_bin_reader(tmp_path):
    from pwclonet_pylidarslam_b200 import scan_input
    import pytest
    raw = syn.make_raw_scan(9)[:1234]
    p = tmp_path / "000000.bin"
    raw.tofile(p)
    got = scan_input.load_kitti_bin(str(p))
    assert got.dtype == np.float32 and got.shape == (1234, 4)
    np.testing.assert_array_equal(got, raw)
    (tmp_path / "bad.bin").write_bytes(b"\x00" * 20)
    with pytest.raises(RuntimeError, match="whole number"):
        scan_input.load_kitti_bin(str(tmp_path / "bad.bin"))
    with pytest.raises(RuntimeError, match="no CPU path"):
        scan_input.prepare_pairs_from_files([str(p)], [str(p)], syn.KITTI_TR, 64, 0, device="cpu")
