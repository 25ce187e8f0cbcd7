"""GPU parity of the stand-alone operators through the C ABI.

Checker = oracle/pointnet2_cpu.c (CPU restatement) and, when oracle/_ref holds it, the reference's
own CUDA kernels recompiled for sm_100a.  Index outputs must be bit-exact."""
import numpy as np
import pytest
import torch

from oracle import cpu_ops
from pwclonet_pylidarslam_b200 import _ext, synthetic

pytestmark = pytest.mark.gpu


def _rand_cloud(rng, B, N, scale=10.0):
    return (rng.standard_normal((B, N, 3)) * np.array([scale, scale * 0.1, scale])).astype(np.float32)


def _dev(a, cuda):
    return torch.from_numpy(np.ascontiguousarray(a)).to(cuda)


FPS_CASES = [
    # (B, N, m, kind)
    (2, 8192, 2048, "lidar"), (3, 8192, 2048, "rand"), (2, 2048, 1024, "rand"), (2, 1024, 256, "rand"),
    (3, 256, 64, "rand"), (2, 64, 16, "rand"), (1, 1, 1, "rand"), (2, 3000, 700, "rand"), (1, 600, 600, "rand"),
    (2, 1000, 300, "dups"), (2, 2048, 512, "origin"), (1, 4096, 512, "grid"), (1, 16384, 1024, "rand"),
    (1, 20000, 300, "rand"),
]


def _fps_input(kind, B, N, seed):
    rng = np.random.default_rng(seed)
    if kind == "lidar":
        return np.stack([synthetic.make_pair(synthetic.SEED_BASE + 100 + i, N)["pc1"] for i in range(B)])
    x = _rand_cloud(rng, B, N)
    if kind == "dups":          # exact duplicates -> exact ties in the arg-max (dataset padding does this)
        x[:, N // 2:] = x[:, : N - N // 2]
    elif kind == "origin":      # points inside the |p|^2 <= 1e-3 ball, incl. index 0
        x[:, ::7] *= 1e-3
        x[:, 0] = 0.0
    elif kind == "grid":        # integer lattice: massive exact distance ties
        g = np.stack(np.meshgrid(np.arange(16), np.arange(16), np.arange(16), indexing="ij"), -1).reshape(-1, 3)
        x = np.broadcast_to(g[None].astype(np.float32), (B, N, 3)).copy() + 1.0
    return x


@pytest.mark.parametrize("B,N,m,kind", FPS_CASES)
def test_fps_bit_exact(cuda, ref_ext, B, N, m, kind):
    x = _fps_input(kind, B, N, seed=N + m)
    want = cpu_ops.fps(x, m, origin_skip=True, thread_cap=512)
    got = _ext.furthest_point_sampling(_dev(x, cuda), m).cpu().numpy()
    assert got.dtype == np.int32 and got.shape == (B, m)
    np.testing.assert_array_equal(got, want)
    if ref_ext is not None:
        ref = ref_ext.furthest_point_sampling(_dev(x, cuda), m).cpu().numpy()
        np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("B,N,m,kind", [(3, 8192, 2048, "lidar"), (2, 8192, 2048, "dups"), (1, 4096, 1024, "grid"), (2, 2048, 1024, "origin"),
                                        (2, 2048, 1024, "rand"), (3, 1024, 256, "dups"), (4, 256, 64, "rand"), (2, 300, 299, "dups"),
                                        (1, 5000, 4999, "rand"), (2, 64, 64, "rand")])
def test_fps_double_pick_equals_single_pick(cuda, monkeypatch, B, N, m, kind):
    """PWCLO_FPS_PICK=2 selects up to two samples per round (exact: the second candidate is taken only when the first
    cannot lower its running minimum; opt-in, measured slower).  Same indices as the default one-sample rounds, and both
    equal the C restatement of the reference kernel -- duplicates, lattice ties, origin-ball points, m == N included."""
    x = _fps_input(kind, B, N, seed=11 * N + m)
    d = _dev(x, cuda)
    one = _ext.furthest_point_sampling(d, m).cpu().numpy()
    monkeypatch.setenv("PWCLO_FPS_PICK", "2")
    two = _ext.furthest_point_sampling(d, m).cpu().numpy()
    np.testing.assert_array_equal(two, one)
    np.testing.assert_array_equal(two, cpu_ops.fps(x, m, origin_skip=True, thread_cap=512))


PREFIX_CASES = [
    # (B, N, kind, cluster): the pyramid chain N -> 2048|N/4 -> /2 -> /4 -> /4 of PW/pwclo_net.py:66-69 on every input kind
    (3, 8192, "lidar", "0"), (2, 8192, "rand", "0"), (2, 8192, "dups", "0"), (2, 8192, "origin", "0"), (1, 4096, "grid", "0"),
    (2, 4096, "rand", "0"), (2, 16384, "lidar", "1"), (2, 16384, "dups", "1"), (2, 4096, "dups", "2"), (2, 4096, "rand", "2"),
    (2, 700, "rand", "0"), (2, 700, "dups", "0"),
]


@pytest.mark.parametrize("B,N,kind,cluster", PREFIX_CASES)
def test_fps_prefix_chain_bit_exact(cuda, monkeypatch, B, N, kind, cluster):
    """pwclo_furthest_point_sampling_prefix along a pyramid: every level equals the plain sampling of the same cloud
    (C oracle) bit for bit, whether the level took the 0..m-1 shortcut (tie-free clouds) or really ran (tie flag set);
    tie-free random clouds must take the shortcut (flag 0), duplicated / lattice clouds must raise the flag."""
    monkeypatch.setenv("PWCLO_FPS_CLUSTER", cluster)
    x = _fps_input(kind, B, N, seed=3 * N + len(kind))
    if kind == "rand" and B > 1:
        x[1, N // 2:] = x[1, :N - N // 2]          # a batch that mixes a tie-free cloud with a duplicated one
    cur, tie = _dev(x, cuda), None
    m = max(N // 4, 16)
    ties = []
    for div in (1, 2, 4, 4):
        m = max(m // div, 4)
        got, tie_out = _ext.furthest_point_sampling(cur, m, tie_in=tie, return_tie=True)
        want = cpu_ops.fps(cur.cpu().numpy(), m, origin_skip=True, thread_cap=512)
        np.testing.assert_array_equal(got.cpu().numpy(), want)
        plain = _ext.furthest_point_sampling(cur, m)
        assert torch.equal(plain, got)
        ties.append(tie_out.cpu().numpy().copy())
        cur = torch.gather(cur, 1, got.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
        tie = tie_out
    t0 = ties[0]
    if kind in ("lidar",) or (kind == "rand" and B == 1):
        assert not t0.any(), "a generic cloud raised the tie flag (the shortcut would never be taken)"
    if kind == "rand" and B > 1:
        assert t0[0] == 0 and t0[1] == 1
    if kind in ("dups", "grid"):
        assert t0.all()
    for a, b in zip(ties[:-1], ties[1:]):     # a clear flag stays clear down the chain
        assert not (b & ~a.astype(bool)).any()


def test_fps_prefix_shortcut_is_taken(cuda):
    """with a clear flag the kernel must not sample at all: hand it a cloud that is NOT FPS-ordered together with a
    clear flag and it returns 0..m-1 (which proves levels 2-4 of a tie-free pyramid cost no sampling rounds)"""
    rng = np.random.default_rng(0)
    x = _dev(_rand_cloud(rng, 2, 2048), cuda)
    flag = torch.tensor([0, 1], dtype=torch.int32, device=cuda)
    got = _ext.furthest_point_sampling(x, 256, tie_in=flag).cpu().numpy()
    want = cpu_ops.fps(x.cpu().numpy(), 256, origin_skip=True, thread_cap=512)
    np.testing.assert_array_equal(got[0], np.arange(256))
    np.testing.assert_array_equal(got[1], want[1])


@pytest.mark.parametrize("B,N,m,kind", [(2, 8192, 1024, "rand"), (2, 1500, 400, "dups"), (1, 512, 100, "origin")])
def test_fps_orphan_variant(cuda, B, N, m, kind):
    """/sampling_gpu_copy.cu: 1024-thread tie order, origin test disabled."""
    x = _fps_input(kind, B, N, seed=7)
    want = cpu_ops.fps(x, m, origin_skip=False, thread_cap=1024)
    got = _ext.furthest_point_sampling(_dev(x, cuda), m, flags=2).cpu().numpy()
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("B,N,m,kind,flags", [(3, 16384, 600, "lidar", 1), (2, 12000, 300, "dups", 1), (2, 4096, 500, "grid", 1),
                                              (4, 4096, 256, "origin", 1), (2, 3000, 200, "dups", 2), (1, 16384, 300, "rand", 2)])
def test_fps_cluster_kernel_bit_exact(cuda, monkeypatch, B, N, m, kind, flags):
    """the 8-CTA cluster kernel (default for n > 8192 with few clouds; forced here for the smaller sizes) against the
    C oracle: exact ties (duplicates, lattice), origin-ball points and both tie orders included"""
    monkeypatch.setenv("PWCLO_FPS_CLUSTER", "2")
    x = _fps_input(kind, B, N, seed=N + m + flags)
    want = cpu_ops.fps(x, m, origin_skip=(flags == 1), thread_cap=512 if flags == 1 else 1024)
    got = _ext.furthest_point_sampling(_dev(x, cuda), m, flags=flags).cpu().numpy()
    np.testing.assert_array_equal(got, want)
    monkeypatch.setenv("PWCLO_FPS_CLUSTER", "0")
    one = _ext.furthest_point_sampling(_dev(x, cuda), m, flags=flags).cpu().numpy()
    np.testing.assert_array_equal(one, want)


def test_fps_nested_prefix_property(cuda):
    """FPS of an FPS prefix is the prefix -- size-independent property at full size."""
    x = np.stack([synthetic.make_pair(synthetic.SEED_BASE + 7, 8192)["pc2"]])
    xd = _dev(x, cuda)
    i1 = _ext.furthest_point_sampling(xd, 2048).long()
    l1 = torch.gather(xd, 1, i1[..., None].expand(-1, -1, 3)).contiguous()
    i2 = _ext.furthest_point_sampling(l1, 1024).cpu().numpy()
    np.testing.assert_array_equal(i2[0], np.arange(1024))
    assert len(set(i1[0].tolist())) == 2048


def torch_knn_point(nsample, xyz, new_xyz):
    """restatement of P2/pytorch_utils.py:12-49 (materialising version), used as GPU checker."""
    diff = new_xyz.unsqueeze(2).repeat(1, 1, xyz.shape[1], 1) - xyz.unsqueeze(1).repeat(1, new_xyz.shape[1], 1, 1)
    dist = torch.sqrt(torch.sum(diff ** 2, dim=-1) + 1e-8)
    d, i = torch.topk(dist, nsample, largest=False, dim=-1)
    return d, i.int()


KNN_CASES = [(2, 8192, 2048, 32), (2, 2048, 1024, 32), (2, 1024, 256, 16), (3, 256, 64, 16), (2, 256, 256, 4),
             (2, 1024, 1024, 6), (2, 64, 256, 8), (1, 100, 37, 5), (1, 9000, 300, 32), (2, 33, 10, 32), (1, 20000, 64, 8)]


@pytest.mark.parametrize("B,N,S,K", KNN_CASES)
@pytest.mark.parametrize("order", [0, 1])
def test_knn_vs_oracle_bit_exact(cuda, B, N, S, K, order):
    rng = np.random.default_rng(N * 31 + S)
    xyz = _rand_cloud(rng, B, N)
    q = xyz[:, rng.permutation(N)[:S]] if S <= N else _rand_cloud(rng, B, S)
    if S > 8:
        q = q.copy()
        q[:, ::3] += rng.standard_normal((B, len(range(0, S, 3)), 3)).astype(np.float32) * 0.05
    want_i, want_d = cpu_ops.knn(xyz, q, K, sum_order=order, return_dist=True)
    got_i, got_d = _ext.knn(_dev(xyz, cuda), _dev(q, cuda), K, sum_order=order, return_dist=True)
    np.testing.assert_array_equal(got_d.cpu().numpy().view(np.int32), want_d.view(np.int32))
    np.testing.assert_array_equal(got_i.cpu().numpy(), want_i)


@pytest.mark.parametrize("B,N,S,K", [(2, 8192, 2048, 32), (3, 2048, 1024, 8), (2, 1000, 333, 6), (2, 100, 64, 32), (1, 64, 7, 4),
                                     (2, 16384, 2048, 32), (2, 12000, 1500, 16), (1, 8193, 700, 4), (3, 16384, 300, 8)])
def test_knn_sorted_equals_bruteforce(cuda, B, N, S, K):
    """the sorted-slab search must be bit-identical to the brute-force kernel (same keys, exact pruning); clouds of more
    than 8192 points take the two-half kernel (the halves of the sorted cloud take turns in shared memory)"""
    lid = np.stack([synthetic.make_pair(synthetic.SEED_BASE + 60 + i, max(N, 8192))["pc2"][:N] for i in range(B)])
    rng = np.random.default_rng(3)
    for xyz in (lid, _rand_cloud(rng, B, N), np.round(_rand_cloud(rng, B, N, 3.0))):   # lidar, gaussian, lattice (ties)
        x = _dev(xyz, cuda)
        qs = [x[:, :S].contiguous() + 0.0]
        if N > 8192:      # queries that are not reference points: shifted into the other half, and far outside the cloud
            qs += [(x[:, -S:] + torch.tensor([7.0, 0.3, -5.0], device=cuda)).contiguous(), (x[:, :S] * 1.7 + 40.0).contiguous()]
        for q in qs:
            old = _ext.KNN_SORTED
            try:
                _ext.KNN_SORTED = True
                i1, d1 = _ext.knn(x, q, K, return_dist=True)
                _ext.KNN_SORTED = False
                i0, d0 = _ext.knn(x, q, K, return_dist=True)
            finally:
                _ext.KNN_SORTED = old
            assert torch.equal(d1, d0) and torch.equal(i1, i0)


def test_knn_presort_search_split(cuda):
    """pwclo_knn_presort once + pwclo_knn_search per query set (also on a sub-batch of the workspace records)
    gives the bits of pwclo_knn_sorted; a far-away, rotated query cloud exercises the 3-D pruning bound"""
    import ctypes
    from pwclonet_pylidarslam_b200 import _lib
    L = _lib.lib()
    B, N = 4, 2048
    lid = np.stack([synthetic.make_pair(synthetic.SEED_BASE + 80 + i, 8192)["pc2"][:N] for i in range(B)])
    x = _dev(lid, cuda)
    c, s_ = np.cos(1.0), np.sin(1.0)
    R = torch.tensor([[1, 0, 0], [0, c, -s_], [0, s_, c]], dtype=torch.float32, device=cuda)
    queries = {"subset": x[:, ::2].contiguous(), "rotated": (x[:, ::3] @ R.T + torch.tensor([3.0, -2.0, 5.0], device=cuda)).contiguous()}
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    for K in (6, 32):
        nbytes = L.pwclo_knn_workspace_bytes(B, N, 0)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=cuda)
        _lib.check(L.pwclo_knn_presort(vp(x), B, N, K, vp(ws), nbytes, _lib.stream_ptr()), "presort")
        for name, q in queries.items():
            S = q.shape[1]
            want_i, want_d = _ext.knn(x, q, K, return_dist=True)
            idx = torch.empty(B, S, K, dtype=torch.int32, device=cuda)
            dist = torch.empty(B, S, K, dtype=torch.float32, device=cuda)
            _lib.check(L.pwclo_knn_search(vp(ws), vp(q), B, N, S, K, _ext.KNN_SUM_ORDER, None, None, vp(idx), vp(dist),
                                          _lib.stream_ptr()), "search")
            assert torch.equal(idx, want_i) and torch.equal(dist, want_d), (K, name)
            # second half of the batch through its own workspace records
            rec = L.pwclo_knn_workspace_bytes(1, N, 0)
            h = B // 2
            idx2 = torch.empty(h, S, K, dtype=torch.int32, device=cuda)
            _lib.check(L.pwclo_knn_search(vp(ws[h * rec:]), vp(q[h:].contiguous()), h, N, S, K, _ext.KNN_SUM_ORDER, None, None,
                                          vp(idx2), None, _lib.stream_ptr()), "search half")
            assert torch.equal(idx2, want_i[h:]), (K, name)
            # and against brute force
            old = _ext.KNN_SORTED
            try:
                _ext.KNN_SORTED = False
                bi, bd = _ext.knn(x, q, K, return_dist=True)
            finally:
                _ext.KNN_SORTED = old
            assert torch.equal(bi, want_i) and torch.equal(bd, want_d), (K, name)


def test_knn_query_order_variant(cuda, monkeypatch):
    """optional Morton query ordering + previous-query bound (PWCLO_KNN_QORDER=1): same bits"""
    lid = np.stack([synthetic.make_pair(synthetic.SEED_BASE + 70 + i, 8192)["pc1"] for i in range(2)])
    x = _dev(lid, cuda)
    q = x[:, ::4].contiguous()
    i0, d0 = _ext.knn(x, q, 16, return_dist=True)
    monkeypatch.setenv("PWCLO_KNN_QORDER", "1")
    i1, d1 = _ext.knn(x, q, 16, return_dist=True)
    assert torch.equal(i0, i1) and torch.equal(d0, d1)


def _tie_aware_equal(got_i, ref_i, ref_d):
    """torch.topk leaves the order among equal distances unspecified: indices must agree wherever
    the reference distance is unique inside the row, and as sets inside groups of equal distance."""
    mism = got_i != ref_i
    n_tie = 0
    for b, s in zip(*np.nonzero(mism.any(-1))):
        d = ref_d[b, s]
        for v in np.unique(d[mism[b, s]]):
            grp = d == v
            # a tie with the (k+1)-th candidate cannot be seen from ref_d alone; allow one swapped index
            a, r = set(got_i[b, s][grp].tolist()), set(ref_i[b, s][grp].tolist())
            assert len(a - r) <= 1 and (grp.sum() > 1 or grp[-1]), (b, s, got_i[b, s], ref_i[b, s], d)
            n_tie += 1
    return n_tie


@pytest.mark.parametrize("B,N,S,K", [(2, 8192, 2048, 32), (2, 2048, 1024, 32), (4, 1024, 256, 16), (2, 1024, 1024, 4)])
def test_knn_vs_torch_cuda(cuda, B, N, S, K):
    """against the reference's own PyTorch formulation executed on the same GPU."""
    xs = np.stack([synthetic.make_pair(synthetic.SEED_BASE + 50 + i, 8192)["pc1"][:N] for i in range(B)])
    xyz = _dev(xs, cuda)
    q = xyz[:, :S].contiguous()
    ref_d, ref_i = torch_knn_point(K, xyz, q)
    got_i, got_d = _ext.knn(xyz, q, K, return_dist=True)
    np.testing.assert_array_equal(got_d.cpu().numpy().view(np.int32), ref_d.cpu().numpy().view(np.int32))
    n_tie = _tie_aware_equal(got_i.cpu().numpy(), ref_i.cpu().numpy(), ref_d.cpu().numpy())
    print(f"knn {B}x{S}x{N} k={K}: rows with tie-order differences: {n_tie}")


def test_knn_fused_warp(cuda):
    rng = np.random.default_rng(5)
    xyz = _rand_cloud(rng, 3, 1024)
    q = _rand_cloud(rng, 3, 256)
    qt = np.concatenate([rng.standard_normal((3, 4)), rng.standard_normal((3, 3))], 1).astype(np.float32)
    qt[:, :4] /= np.linalg.norm(qt[:, :4], axis=1, keepdims=True)
    idx, warped = _ext.knn(_dev(xyz, cuda), _dev(q, cuda), 6, sum_order=0, warp_qt=_dev(qt, cuda), return_warped=True)
    w = warped.cpu().numpy()
    # rotation check in float64
    def rot(qv, p):
        w0, x, y, z = qv
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w0), 2 * (x * z + y * w0)],
                      [2 * (x * y + z * w0), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w0)],
                      [2 * (x * z - y * w0), 2 * (y * z + x * w0), 1 - 2 * (x * x + y * y)]])
        return p @ R.T
    for b in range(3):
        np.testing.assert_allclose(w[b], rot(qt[b, :4].astype(np.float64), q[b].astype(np.float64)) + qt[b, 4:], rtol=0, atol=2e-5)
    want = cpu_ops.knn(xyz, w, 6, sum_order=0)
    np.testing.assert_array_equal(idx.cpu().numpy(), want)


GROUP_CASES = [(2, 3, 8192, 2048, 32), (2, 16, 2048, 1024, 32), (2, 64, 256, 256, 4), (1, 5, 1000, 77, 3),
               (2, 67, 1024, 256, 16), (1, 130, 64, 64, 8), (1, 32, 16384, 512, 16)]


@pytest.mark.parametrize("B,C,N,S,K", GROUP_CASES)
def test_group_points_and_grad(cuda, ref_ext, B, C, N, S, K):
    rng = np.random.default_rng(C * N)
    pts = rng.standard_normal((B, C, N)).astype(np.float32)
    idx = rng.integers(0, N, size=(B, S, K)).astype(np.int32)
    got = _ext.group_points(_dev(pts, cuda), _dev(idx, cuda))
    np.testing.assert_array_equal(got.cpu().numpy(), cpu_ops.group_points(pts, idx))
    if ref_ext is not None:
        assert torch.equal(got, ref_ext.group_points(_dev(pts, cuda), _dev(idx, cuda)))
    go = rng.standard_normal((B, C, S, K)).astype(np.float32)
    gg = _ext.group_points_grad(_dev(go, cuda), _dev(idx, cuda), N).cpu().numpy()
    np.testing.assert_allclose(gg, cpu_ops.group_points_grad(go, idx, N), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("B,C,N,M", [(2, 3, 8192, 2048), (3, 3, 256, 64), (1, 7, 1000, 333)])
def test_gather_points_and_grad(cuda, ref_ext, B, C, N, M):
    rng = np.random.default_rng(N + M)
    pts = rng.standard_normal((B, C, N)).astype(np.float32)
    idx = rng.integers(0, N, size=(B, M)).astype(np.int32)
    got = _ext.gather_points(_dev(pts, cuda), _dev(idx, cuda))
    np.testing.assert_array_equal(got.cpu().numpy(), cpu_ops.gather_points(pts, idx))
    if ref_ext is not None:
        assert torch.equal(got, ref_ext.gather_points(_dev(pts, cuda), _dev(idx, cuda)))
    go = rng.standard_normal((B, C, M)).astype(np.float32)
    gg = _ext.gather_points_grad(_dev(go, cuda), _dev(idx, cuda), N).cpu().numpy()
    np.testing.assert_allclose(gg, cpu_ops.gather_points_grad(go, idx, N), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("B,n,m,r,ns", [(2, 2048, 512, 1.5, 32), (2, 1000, 100, 0.2, 16), (1, 64, 64, 100.0, 8), (1, 500, 50, 1e-4, 4)])
def test_ball_query(cuda, ref_ext, B, n, m, r, ns):
    rng = np.random.default_rng(n + m)
    xyz = _rand_cloud(rng, B, n, scale=3.0)
    q = xyz[:, :m].copy() + 0.01
    got = _ext.ball_query(_dev(q, cuda), _dev(xyz, cuda), r, ns)
    np.testing.assert_array_equal(got.cpu().numpy(), cpu_ops.ball_query(q, xyz, r, ns))
    if ref_ext is not None:
        assert torch.equal(got, ref_ext.ball_query(_dev(q, cuda), _dev(xyz, cuda), r, ns))


@pytest.mark.parametrize("B,n,m,c", [(2, 1024, 256, 16), (1, 333, 77, 5), (1, 10, 2, 3)])
def test_three_nn_interpolate(cuda, ref_ext, B, n, m, c):
    rng = np.random.default_rng(n * m)
    unknown = _rand_cloud(rng, B, n, 3.0)
    known = _rand_cloud(rng, B, m, 3.0)
    if m > 4:
        known[:, 3] = known[:, 1]   # exact distance ties -> lowest index must win
    d2, idx = _ext.three_nn(_dev(unknown, cuda), _dev(known, cuda))
    wd2, widx = cpu_ops.three_nn(unknown, known)
    np.testing.assert_array_equal(idx.cpu().numpy(), widx)
    np.testing.assert_array_equal(d2.cpu().numpy(), wd2)
    if ref_ext is not None and m >= 3:
        rd2, ridx = ref_ext.three_nn(_dev(unknown, cuda), _dev(known, cuda))
        assert torch.equal(idx, ridx) and torch.equal(d2, rd2)
    if m < 3:
        return
    feats = rng.standard_normal((B, c, m)).astype(np.float32)
    w = rng.random((B, n, 3)).astype(np.float32)
    out = _ext.three_interpolate(_dev(feats, cuda), idx, _dev(w, cuda))
    np.testing.assert_array_equal(out.cpu().numpy(), cpu_ops.three_interpolate(feats, widx, w))
    if ref_ext is not None:
        assert torch.equal(out, ref_ext.three_interpolate(_dev(feats, cuda), idx, _dev(w, cuda)))
    go = rng.standard_normal((B, c, n)).astype(np.float32)
    gg = _ext.three_interpolate_grad(_dev(go, cuda), idx, _dev(w, cuda), m).cpu().numpy()
    np.testing.assert_allclose(gg, cpu_ops.three_interpolate_grad(go, widx, w, m), rtol=1e-5, atol=1e-5)


def test_errors_raise_not_exit(cuda):
    x = torch.zeros(1, 16, 3, device=cuda)
    with pytest.raises(RuntimeError):
        _ext.furthest_point_sampling(x.cpu(), 4)
    with pytest.raises(RuntimeError):
        _ext.furthest_point_sampling(x.double(), 4)
    with pytest.raises(RuntimeError):
        _ext.group_points(x.transpose(1, 2), torch.zeros(1, 2, 2, dtype=torch.int32, device=cuda))
    with pytest.raises(RuntimeError):
        _ext.knn(x, x, 64)
