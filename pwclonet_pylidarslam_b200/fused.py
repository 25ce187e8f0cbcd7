"""Fused inference engine for PWCLONet (eval mode): BatchNorm folded into the 1x1 convolutions,
one sm_100a kernel per layer, neighbour search with the pose warp fused in, both frames of the
siamese pyramid batched into one launch per level.  Everything is driven through the C ABI of
libpwclo_b200.so (include/pwclo_b200.h); torch only owns device memory and the stream.

Layer-by-layer correspondence with the reference (PW = slam/models/PWCLONet,
P2 = .../pointnet2_ops):
    set_conv          P2/pointnet2_modules.py:179-245      (FPS -> gather -> kNN -> grouped MLP -> max)
    set_upconv        P2/pointnet2_modules.py:459-515
    cost_volume       PW/costvolume.py:63-190
    flow_predictor    PW/flowpredictor.py:53-84
    pose_head         PW/pose_calculator.py:47-87 + PW/pose_warp_refinement.py:120-148
    forward           PW/pwclo_net.py:109-207
"""
import ctypes

import numpy as np
import torch

import contextlib
import os

from . import _ext, _lib, tc_pack

BN_EPS = 1e-5


LayerT = _lib.LayerT
_vp, _i = ctypes.c_void_p, ctypes.c_int


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


# ------------------------------------------------------------------------------------------------
# host-side weight preparation
# ------------------------------------------------------------------------------------------------
def fold_conv_bn(sd, prefix):
    """[Conv2d 1x1 (no bias) -> BatchNorm2d(eval)] -> (W [cout,cin], b [cout]) in float64
    (P2/pytorch_utils.py:114-167; BN eps = nn.BatchNorm2d default)."""
    w = sd[prefix + ".conv.weight"].double().reshape(sd[prefix + ".conv.weight"].shape[0], -1)
    g, beta = sd[prefix + ".bn.bn.weight"].double(), sd[prefix + ".bn.bn.bias"].double()
    mu, var = sd[prefix + ".bn.bn.running_mean"].double(), sd[prefix + ".bn.bn.running_var"].double()
    s = g / torch.sqrt(var + BN_EPS)
    return w * s[:, None], beta - mu * s


def pack_layer(W, b, k_order=None, k4=None):
    """W [cout,cin] -> the C-ABI layout [cout/CB][k4][CB] (CB = min(cout,64)), see include/pwclo_b200.h.
    k_order: list of source columns (or -1 for a zero row) giving the kernel's internal input order."""
    W = W.detach().double().cpu().numpy() if torch.is_tensor(W) else np.asarray(W, np.float64)
    b = b.detach().double().cpu().numpy() if torch.is_tensor(b) else np.asarray(b, np.float64)
    cout, cin = W.shape
    if k_order is None:
        k_order = list(range(cin))
    kk = len(k_order)
    k4 = (kk + 3) & ~3 if k4 is None else k4
    assert k4 >= kk and k4 % 4 == 0
    Wt = np.zeros((k4, cout), np.float64)
    for dst, src in enumerate(k_order):
        if src >= 0:
            Wt[dst] = W[:, src]
    cb = min(cout, 64)
    assert cout % cb == 0
    blocked = Wt.reshape(k4, cout // cb, cb).transpose(1, 0, 2)
    return np.ascontiguousarray(blocked, np.float32).reshape(-1), b.astype(np.float32), cin, cout


class WeightArena:
    """All folded weights in ONE device buffer (3.1 MB), every tensor 256-byte aligned."""

    def __init__(self):
        self.chunks, self.size, self.pending = [], 0, []

    def add(self, arr):
        off = self.size
        self.chunks.append((off, np.ascontiguousarray(arr, np.float32).reshape(-1)))
        self.size = (off + arr.size * 4 + 255) & ~255
        return off

    def add_layer(self, packed):
        w, b, cin, cout = packed
        rec = {"w_off": self.add(w), "b_off": self.add(b), "cin": cin, "cout": cout}
        self.pending.append(rec)
        return rec

    def finalize(self, device):
        host = np.zeros(self.size // 4, np.float32)
        for off, a in self.chunks:
            host[off // 4: off // 4 + a.size] = a
        self.buf = torch.from_numpy(host).to(device)
        self.base = self.buf.data_ptr()
        return self

    def ptr(self, off):
        return self.base + off

    def layers(self, recs):
        arr = (LayerT * len(recs))()
        for i, r in enumerate(recs):
            arr[i].w, arr[i].b, arr[i].cin, arr[i].cout = self.ptr(r["w_off"]), self.ptr(r["b_off"]), r["cin"], r["cout"]
        return arr


def _n_layers(sd, prefix):
    n = 0
    while f"{prefix}.layer{n}.conv.weight" in sd:
        n += 1
    return n


class FusedPWCLONet:
    def __init__(self, net):
        sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}   # fold BN on the host, once
        dev = next(net.parameters()).device
        if dev.type != "cuda":
            raise _lib.PwcloError("the fused PWCLO-Net engine needs the parameters on a CUDA device (no CPU fallback)")
        self.device = dev
        A = WeightArena()
        self.A = A
        self.recs = {}

        self.tc_recs = {}

        def mlp(prefix, first_order=None, first_k4=None, tc_first_order=None):
            recs, tcr = [], []
            for i in range(_n_layers(sd, prefix)):
                W, b = fold_conv_bn(sd, f"{prefix}.layer{i}")
                recs.append(A.add_layer(pack_layer(W, b, first_order if i == 0 else None, first_k4 if i == 0 else None)))
                if W.shape[0] in (64, 128):       # tensor-core packing (tcgen05 kernels need cout 64 / 128)
                    Wn = W.numpy()
                    if i == 0 and tc_first_order is not None:
                        Wi = np.zeros((Wn.shape[0], len(tc_first_order)))
                        for dst, src in enumerate(tc_first_order):
                            if src >= 0:
                                Wi[:, dst] = Wn[:, src]
                        Wn = Wi
                    tcr.append({"w_off": A.add(tc_pack.pack_tc2(Wn)), "b_off": A.add(b.float().numpy()),
                                "cin": W.shape[1], "cout": W.shape[0]})
            self.recs[prefix] = recs
            if len(tcr) == len(recs):
                self.tc_recs[prefix] = tcr

        def sa_order(cin):  # reference (xyz_diff(3), feat(C)) -> internal (feat(C), xyz_diff(3))
            return list(range(3, cin)) + [0, 1, 2]

        for name in ("psa_1", "psa_2", "psa_3", "psa_4", "flow_feature_encoding"):
            cin = sd[f"{name}.mlp_module.layer0.conv.weight"].shape[1]
            mlp(f"{name}.mlp_module", sa_order(cin), tc_first_order=sa_order(cin) + [-1] * 13)

        def cost_volume(prefix):
            cin = sd[f"{prefix}.mlp_convs.layer0.conv.weight"].shape[1]
            order = list(range(10)) + [-1, -1] + list(range(10, cin))     # (geo(10), 0, 0, f1, f2)
            geo_tc = list(range(10)) + [-1] * 6
            mlp(f"{prefix}.mlp_convs", order, tc_first_order=list(range(10, cin)) + geo_tc)   # (f1, f2, geo(10), 0 x6)
            mlp(f"{prefix}.mlp_conv_xyz_1", tc_first_order=geo_tc)
            mlp(f"{prefix}.mlp_conv_xyz_2", tc_first_order=geo_tc)
            mlp(f"{prefix}.mlp2_convs")
            mlp(f"{prefix}.mlp3_convs")

        def pose_calc(prefix):
            r = {}
            for nm in ("conv1d_q_t", "conv1d_q", "conv1d_t"):
                w = sd[f"{prefix}.{nm}.conv.weight"].float().reshape(sd[f"{prefix}.{nm}.conv.weight"].shape[0], -1)
                r[nm + ".w"] = A.add(w.cpu().numpy())
                r[nm + ".b"] = A.add(sd[f"{prefix}.{nm}.conv.bias"].float().cpu().numpy())
            self.recs[prefix] = r

        cost_volume("cost_volume")
        mlp("l4_flow_predictor.mlp_convs")
        pose_calc("pose_calculator_4")
        for l in (3, 2, 1):
            p = f"pose_warp_refinement_{l}"
            for up in ("setupconv_features", "setupconv_mask"):
                cin_up = sd[f"{p}.{up}.mlp.layer0.conv.weight"].shape[1]
                mlp(f"{p}.{up}.mlp", tc_first_order=list(range(cin_up)) + [-1] * 13)   # (features, xyz_diff) + pad to 16
                mlp(f"{p}.{up}.post_mlp")
            cost_volume(f"{p}.cost_volume")
            mlp(f"{p}.flow_predictor_features.mlp_convs")
            if l != 1:
                mlp(f"{p}.flow_predictor_mask.mlp_convs")
            pose_calc(f"{p}.pose_calculator")
        A.finalize(dev)
        self.L = {k: A.layers(v) for k, v in self.recs.items() if isinstance(v, list)}
        self.LT = {k: A.layers(v) for k, v in self.tc_recs.items()}
        self.use_tc = os.environ.get("PWCLO_TC", "1") != "0"
        # FPS prefix shortcut (pwclo_furthest_point_sampling_prefix): OFF by default.  Measured on the B200
        # (tools/bench_fps_prefix.py, profiles/round2_fps_prefix.json): tracking arg-max ties costs +40 % on the level-1
        # sampling (0.89 -> 1.26 ms: the check sits on the serial round chain), ~6 % of the synthetic LiDAR clouds do see an
        # exact fp32 tie of two running minima in 2047 rounds, and a launch lasts as long as its slowest cloud -- so with
        # 128 clouds per launch levels 2-4 (0.39 ms) are never skipped.  It pays only for single tie-free clouds.
        self.fps_prefix = os.environ.get("PWCLO_FPS_PREFIX", "0") == "1"
        self._sorted = {}
        self._graphs = {}        # (B, N) -> captured whole-forward CUDA graph (forward_graphed)
        self.use_graph = os.environ.get("PWCLO_INFER_GRAPH", "1") != "0"
        self.lib = _lib.lib()
        self.launches = 0
        self.verbose_timeline = False
        self.timeline = None     # when a list: (kernel name, start event, end event) per launch
        # Optional (PWCLO_OVERLAP=1): everything that depends on coordinates only (the FPS chain, the gathers and every
        # kNN search except the two per level against the pose-warped cloud) on a second stream, ahead of the layer
        # kernels that consume it.  Measured on the B200: 7.49 ms per 64 pairs against 7.31 ms on one stream -- every
        # kernel here already fills the machine (the kNN CTAs take 1024 threads x 63 registers, nothing co-resides),
        # so the streams only interleave whole kernels and the geometry chain, which is the critical path, loses SMs
        # to the layers.  Off by default.
        # ... but with FEW clouds (a sharded batch of 8 pairs, one odometry pair) the FPS chain occupies 2B SMs for
        # ~1.3 ms while everything else waits: there the second stream lets the level-l set conv / kNN run beside the
        # level-(l+1) sampling.  PWCLO_OVERLAP = 1 / 0 forces it on / off; default: on when 2B <= OVERLAP_MAX_CLOUDS.
        ov = os.environ.get("PWCLO_OVERLAP", "")
        self.overlap_mode = {"1": True, "0": False}.get(ov)
        self.overlap = False
        self._side = torch.cuda.Stream(device=dev)
        self._branch = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))   # the two set-upconvs of a level

    # ------------------------------------------------------------------ thin launch helpers
    def _call(self, name, *args, note="", work=(0, 0)):
        """work = (algorithmic HBM bytes, MLP flops) of this launch, recorded with the timeline"""
        if self.timeline is not None:
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
        rc = getattr(self.lib, name)(*args, _lib.stream_ptr())
        _lib.check(rc, name)
        self.launches += 1
        if self.timeline is not None:
            e.record()
            self.timeline.append((name + note, s, e, work))

    def _flops(self, key, rows):
        return 2 * rows * sum(r["cin"] * r["cout"] for r in self.recs[key])

    def _new(self, *shape, dtype=torch.float32):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def fps(self, xyz, m, tie_in=None):
        """-> (idx [B,m], tie [B]): `tie` is the per-cloud flag of pwclo_furthest_point_sampling_prefix; handing it to
        the next level's call lets tie-free clouds return 0..m-1 without their m-1 dependent rounds (same indices)"""
        B, N, _ = xyz.shape
        idx = self._new(B, m, dtype=torch.int32)
        if not self.fps_prefix:
            self._call("pwclo_furthest_point_sampling", _p(xyz), B, N, m, 1, _p(idx),
                       note=f"[B{B} N{N} m{m}]" if self.verbose_timeline else "", work=(4 * B * (3 * N + m), 0))
            return idx, None
        tie = self._new(B, dtype=torch.int32)
        self._call("pwclo_furthest_point_sampling_prefix", _p(xyz), B, N, m, 1, _p(idx), _p(tie_in), _p(tie),
                   note=f"[B{B} N{N} m{m}]" if self.verbose_timeline else "", work=(4 * B * (3 * N + m), 0))
        return idx, tie

    def gather3(self, xyz, idx):
        B, N, _ = xyz.shape
        out = self._new(B, idx.shape[1], 3)
        self._call("pwclo_gather_rows3", _p(xyz), _p(idx), B, N, idx.shape[1], _p(out))
        return out

    def knn(self, xyz, queries, k, warp_qt=None, keep_sorted=False):
        """Exact kNN of `queries` in `xyz`.  keep_sorted: remember the sorted workspace of these reference clouds
        (and of the two halves of the batch) for later searches of the same forward (`_sorted`)."""
        B, N, _ = xyz.shape
        S = queries.shape[1]
        idx = self._new(B, S, k, dtype=torch.int32)
        warped = self._new(B, S, 3) if warp_qt is not None else None
        note = f"[B{B} S{S} N{N} k{k}]" if self.verbose_timeline else ""
        work = (4 * B * (3 * S + 3 * N + S * k), 0)
        ws_bytes = self.lib.pwclo_knn_workspace_bytes(B, N, S) if _ext.KNN_SORTED and N >= _ext.KNN_SORTED_MIN_N else 0
        hit = self._sorted.get((xyz.data_ptr(), B, N)) if ws_bytes else None
        if hit is not None:      # these clouds were sorted earlier in this forward: search only
            self._call("pwclo_knn_search", _p(hit), _p(queries), B, N, S, k, _ext.KNN_SUM_ORDER, _p(warp_qt), _p(warped),
                       _p(idx), None, note=note, work=work)
        elif ws_bytes:
            ws = self._new(ws_bytes, dtype=torch.uint8)
            self._call("pwclo_knn_sorted", _p(xyz), _p(queries), B, N, S, k, _ext.KNN_SUM_ORDER, _p(warp_qt), _p(warped),
                       _p(idx), None, _p(ws), ws_bytes, note=note, work=work)
            if keep_sorted and B % 2 == 0 and not os.environ.get("PWCLO_KNN_QORDER"):
                rec = self.lib.pwclo_knn_workspace_bytes(1, N, 0)           # bytes per cloud record
                h = B // 2
                self._sorted[(xyz.data_ptr(), B, N)] = ws
                self._sorted[(xyz[:h].data_ptr(), h, N)] = ws[:h * rec]
                self._sorted[(xyz[h:].data_ptr(), h, N)] = ws[h * rec:B * rec]
        else:
            self._call("pwclo_knn", _p(xyz), _p(queries), B, N, S, k, _ext.KNN_SUM_ORDER, _p(warp_qt), _p(warped), _p(idx),
                       None, note=note, work=work)
        return (idx, warped) if warp_qt is not None else idx

    def presort(self, xyz, k):
        """sort the clouds `xyz` [B,N,3] for later searches of this forward (pwclo_knn_presort) and remember the workspace
        for the whole batch and for its two halves (frame 1 / frame 2)"""
        B, N, _ = xyz.shape
        if not (_ext.KNN_SORTED and N >= _ext.KNN_SORTED_MIN_N) or B % 2 or os.environ.get("PWCLO_KNN_QORDER"):
            return
        ws_bytes = self.lib.pwclo_knn_workspace_bytes(B, N, 0)
        if not ws_bytes or (xyz.data_ptr(), B, N) in self._sorted:
            return
        ws = self._new(ws_bytes, dtype=torch.uint8)
        self._call("pwclo_knn_presort", _p(xyz), B, N, k, _p(ws), ws_bytes,
                   note=f"[B{B} N{N}]" if self.verbose_timeline else "", work=(4 * B * 3 * N, 0))
        rec = self.lib.pwclo_knn_workspace_bytes(1, N, 0)
        h = B // 2
        self._sorted[(xyz.data_ptr(), B, N)] = ws
        self._sorted[(xyz[:h].data_ptr(), h, N)] = ws[:h * rec]
        self._sorted[(xyz[h:].data_ptr(), h, N)] = ws[h * rec:B * rec]

    def set_conv(self, key, xyz, feats, new_xyz, idx):
        B, N, _ = xyz.shape
        S, K = idx.shape[1], idx.shape[2]
        layers = self.L[key]
        out = self._new(B, S, layers[len(layers) - 1].cout)
        C = feats.shape[2] if feats is not None else 3
        if self.use_tc and key in self.LT and feats is not None and C % 16 == 0:
            tl = self.LT[key]
            self._call("pwclo_set_conv_tc", _p(xyz), _p(feats), _p(new_xyz), _p(idx), B, N, S, K, C, tl, len(tl), _p(out),
                       note=f"[{key} B{B} S{S} K{K} C{C}]" if self.verbose_timeline else "",
                       work=(4 * B * (S * K + S * out.shape[2] + N * (C + 3)), self._flops(key, B * S * K)))
            return out
        self._call("pwclo_set_conv", _p(xyz), _p(feats), _p(new_xyz), _p(idx), B, N, S, K, C, layers, len(layers), _p(out),
                   note=f"[{key} B{B} S{S} K{K} C{C}]" if self.verbose_timeline else "",
                   work=(4 * B * (S * K + S * out.shape[2] + N * (C + 3)), self._flops(key, B * S * K)))
        return out

    def pointwise(self, key, srcs):
        rows = srcs[0].shape[0] * srcs[0].shape[1]
        layers = self.L[key]
        out = self._new(srcs[0].shape[0], srcs[0].shape[1], layers[len(layers) - 1].cout)
        ptrs = (_vp * len(srcs))(*[s.data_ptr() for s in srcs])
        chans = (_i * len(srcs))(*[s.shape[2] for s in srcs])
        if self.use_tc and key in self.LT and all(s.shape[2] % 16 == 0 for s in srcs):
            tl = self.LT[key]
            self._call("pwclo_pointwise_mlp_tc", ptrs, chans, len(srcs), rows, tl, len(tl), _p(out),
                       work=(4 * rows * (sum(s.shape[2] for s in srcs) + out.shape[2]), self._flops(key, rows)))
            return out
        self._call("pwclo_pointwise_mlp", ptrs, chans, len(srcs), rows, layers, len(layers), _p(out),
                   work=(4 * rows * (sum(s.shape[2] for s in srcs) + out.shape[2]), self._flops(key, rows)))
        return out

    def cost_volume(self, prefix, wxyz, f1, xyz2, f2, idx_q, idx_self):
        B, S, _ = wxyz.shape
        N, C = xyz2.shape[1], f1.shape[2]
        e1 = self._new(B, S, 64)
        Kq, Ks = idx_q.shape[2], idx_self.shape[2]
        fl1 = sum(self._flops(prefix + k, B * S * Kq) for k in (".mlp_convs", ".mlp_conv_xyz_1", ".mlp2_convs"))
        fl2 = sum(self._flops(prefix + k, B * S * Ks) for k in (".mlp_conv_xyz_2", ".mlp3_convs"))
        w1 = (4 * B * (S * (3 + C + Kq + 64) + N * (3 + C)), fl1)
        w2 = (4 * B * S * (3 + C + 64 + Ks + 64), fl2)
        if self.use_tc:
            note1 = f"[B{B} S{S} K{idx_q.shape[2]} C{C}]" if self.verbose_timeline else ""
            self._call("pwclo_cost_volume_1_tc", _p(wxyz), _p(f1), _p(xyz2), _p(f2), _p(idx_q), B, S, N, idx_q.shape[2], C,
                       self.LT[prefix + ".mlp_convs"], self.LT[prefix + ".mlp_conv_xyz_1"], self.LT[prefix + ".mlp2_convs"],
                       _p(e1), note=note1, work=w1)
            out = self._new(B, S, 64)
            self._call("pwclo_cost_volume_2_tc", _p(wxyz), _p(f1), _p(e1), _p(idx_self), B, S, idx_self.shape[2], C,
                       self.LT[prefix + ".mlp_conv_xyz_2"], self.LT[prefix + ".mlp3_convs"], _p(out),
                       note=f"[B{B} S{S} C{C}]" if self.verbose_timeline else "", work=w2)
            return out, e1
        self._call("pwclo_cost_volume_1", _p(wxyz), _p(f1), _p(xyz2), _p(f2), _p(idx_q), B, S, N, idx_q.shape[2], C,
                   self.L[prefix + ".mlp_convs"], self.L[prefix + ".mlp_conv_xyz_1"], self.L[prefix + ".mlp2_convs"], _p(e1),
                   note=f"[B{B} S{S} K{idx_q.shape[2]} C{C}]" if self.verbose_timeline else "", work=w1)
        out = self._new(B, S, 64)
        self._call("pwclo_cost_volume_2", _p(wxyz), _p(f1), _p(e1), _p(idx_self), B, S, idx_self.shape[2], C,
                   self.L[prefix + ".mlp_conv_xyz_2"], self.L[prefix + ".mlp3_convs"], _p(out),
                   note=f"[B{B} S{S} C{C}]" if self.verbose_timeline else "", work=w2)
        return out, e1

    def pose_head(self, prefix, emb, mask, coarse_qt, pose_params, level):
        B, S, _ = emb.shape
        r = self.recs[prefix]
        qt = self._new(B, 7)
        P = self.A.ptr
        self._call("pwclo_pose_head", _p(emb), _p(mask), B, S, P(r["conv1d_q_t.w"]), P(r["conv1d_q_t.b"]),
                   P(r["conv1d_q.w"]), P(r["conv1d_q.b"]), P(r["conv1d_t.w"]), P(r["conv1d_t.b"]), _p(coarse_qt), _p(qt),
                   _p(pose_params), level)
        return qt

    def to_point_major(self, x):
        """[B,C,N] -> [B,N,C]"""
        B, C, N = x.shape
        out = self._new(B, N, C)
        self._call("pwclo_transpose", _p(x.contiguous()), B, C, N, 1, _p(out))
        return out

    # ------------------------------------------------------------------ the network
    LEVELS = ((2048, 32), (1024, 32), (256, 16), (64, 16))

    # ------------------------------------------------------------------ the whole forward as ONE CUDA graph
    MAX_GRAPHS = 32     # (shape, slot) records: six forwards in flight (sharding.ForwardStreams) x a few shapes
    OVERLAP_MAX_CLOUDS = 48

    def forward_graphed(self, xyz_f1, xyz_f2, slot=0):
        """forward() replayed from a CUDA graph captured once per input shape (every shape on the path is static:
        the launch list of a forward depends on (B, N) only).  The reference pays ~550 launches and a forced
        device->host sync per forward (PW/pwclo_net.py:186-193); the fused forward is ~55 launches whose host cost
        (ctypes + ~60 torch.empty) is what bounds batches of a few pairs -- the sharded batch at 8 GPUs, the odometry
        adapter at one pair.  Inputs are copied into the graph's static buffers (this replaces forward()'s torch.cat),
        outputs are cloned out of them, so results stay valid across calls like any torch result.
        `slot`: forwards that are in flight at the same time on different streams (sharding.ForwardStreams) need their
        own static buffers: one captured graph per (shape, slot)."""
        if (not self.use_graph or self.timeline is not None or not xyz_f1.is_cuda
                or xyz_f1.dtype != torch.float32 or xyz_f2.dtype != torch.float32 or xyz_f1.shape != xyz_f2.shape):
            return self.forward(xyz_f1, xyz_f2)
        key = (xyz_f1.shape[0], xyz_f1.shape[2]) if slot == 0 else (xyz_f1.shape[0], xyz_f1.shape[2], slot)
        rec = self._graphs.get(key)
        if rec is None:
            rec = self._capture(xyz_f1, xyz_f2)
            if len(self._graphs) >= self.MAX_GRAPHS:
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = rec
        if rec is False:                       # capture failed for this shape (reported once): eager launches
            return self.forward(xyz_f1, xyz_f2)
        B = key[0]
        with torch.cuda.device(self.device):
            rec["in"][:B].copy_(xyz_f1, non_blocking=True)
            rec["in"][B:].copy_(xyz_f2, non_blocking=True)
            rec["graph"].replay()
            self.launches += rec["launches"]
            return tuple(o.clone() for o in rec["out"])

    def _capture(self, xyz_f1, xyz_f2):
        B, _, N = xyz_f1.shape
        with torch.cuda.device(self.device):
            cur = torch.cuda.current_stream(self.device)
            static_in = torch.cat((xyz_f1, xyz_f2), dim=0).contiguous().clone()
            try:
                side = torch.cuda.Stream(device=self.device)
                side.wait_stream(cur)
                with torch.cuda.stream(side):             # one eager pass: loads every kernel, sets the smem attributes
                    self.forward(None, None, _cat=static_in)
                cur.wait_stream(side)
                torch.cuda.synchronize(self.device)
                graph = torch.cuda.CUDAGraph()
                n0 = self.launches
                with torch.cuda.graph(graph):
                    out = self.forward(None, None, _cat=static_in)
                n = self.launches - n0
                self.launches = n0
                return {"in": static_in, "graph": graph, "out": out, "launches": n}
            except Exception as e:      # the same kernels still run, one launch at a time
                import warnings
                warnings.warn(f"CUDA-graph capture of the fused forward failed for B={B}, N={N} ({type(e).__name__}: {e}); "
                              "this shape runs with eager launches")
                torch.cuda.synchronize(self.device)
                return False

    def forward(self, xyz_f1, xyz_f2, trace=None, _cat=None):
        """xyz_f1, xyz_f2: [B,3,N] fp32 CUDA -> (pose_params [B,4,7], embedding_mask_1 [B,64,2048] view,
        new_xyz_f1_1 [B,2048,3]).  _cat: the two frames already stacked as [2B,3,N] (forward_graphed)."""
        B = xyz_f1.shape[0] if _cat is None else _cat.shape[0] // 2
        self.overlap = (self.overlap_mode if self.overlap_mode is not None else 2 * B <= self.OVERLAP_MAX_CLOUDS) \
            and self.timeline is None
        self._sorted = {}          # (data_ptr, clouds, points) -> sorted kNN workspace, valid for this forward only
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream(self.device)
            side = self._side if self.overlap else None
            geo = contextlib.nullcontext() if side is None else torch.cuda.stream(side)

            def mark():
                """event after the geometry work issued so far (None when everything runs on one stream)"""
                if side is None:
                    return None
                ev = torch.cuda.Event()
                ev.record(side)
                return ev

            def need(ev):
                if ev is not None:
                    main.wait_event(ev)

            # siamese pyramid: both frames share the weights -> one batch of 2B clouds per level
            xyz = self.to_point_major(torch.cat((xyz_f1, xyz_f2), dim=0).float() if _cat is None else _cat)
            if side is not None:
                side.wait_stream(main)      # inputs are ready; also orders this forward's geometry after the previous forward
            # Few-cloud mode: the side stream carries ONLY the dependent sampling chain (FPS -> gather per level, ~1.3 ms
            # on 2B SMs); the neighbour searches and set convs of level l run on the main stream as soon as level l's
            # centres exist, i.e. beside the sampling of levels l+1...  With many clouds everything is one stream.
            xs, fidxs, lvl_ev = [xyz], [], []
            with geo:
                tie = None
                for l, (npoint, k) in enumerate(self.LEVELS):
                    fidx, tie = self.fps(xs[-1], npoint, tie)
                    xs.append(self.gather3(xs[-1], fidx))
                    fidxs.append(fidx)
                    lvl_ev.append(mark())
            X1 = [None] + [x[:B] for x in xs[1:]]
            X2 = [None] + [x[B:] for x in xs[1:]]
            fs, lvl_idx, up_idx = [None], [], {}
            for l, (npoint, k) in enumerate(self.LEVELS):
                need(lvl_ev[l])
                idx = self.knn(xs[l], xs[l + 1], k, keep_sorted=l >= 1)          # levels 1-3 are searched again below
                lvl_idx.append((fidxs[l], idx))
                fs.append(self.set_conv(f"psa_{l + 1}.mlp_module", xs[l], fs[-1], xs[l + 1], idx))
                if l < 3:       # the centres of this level are the reference cloud of several later searches: sort them once, now
                    self.presort(xs[l + 1], self.LEVELS[l + 1][1])
                if l == 2:      # coordinate-only searches whose inputs exist now: coarse cost volume (level 3), upconvs 2 and 1
                    idx_q3 = self.knn(X2[3], X1[3], 32)
                    idx_s3 = self.knn(X1[3], X1[3], 4)
                    up_idx[2] = self.knn(X1[3], X1[2], 8)
                    up_idx[1] = self.knn(X1[2], X1[1], 8)
            up_idx[3] = self.knn(X1[4], X1[3], 8)
            up_ev = {3: None, 2: None, 1: None}
            F1 = [None] + [f[:B] for f in fs[1:]]
            F2 = [None] + [f[B:] for f in fs[1:]]
            pose = self._new(B, 4, 7)

            # coarse cost volume at level 3 + flow feature encoding (PW/pwclo_net.py:162-167)
            emb, _ = self.cost_volume("cost_volume", X1[3], F1[3], X2[3], F2[3], idx_q3, idx_s3)
            # flow_feature_encoding re-runs FPS + kNN on xyz1 of level 3: identical to psa_4's (frame 1)
            emb4 = self.set_conv("flow_feature_encoding.mlp_module", X1[3], emb, X1[4], lvl_idx[3][1][:B])
            mask4 = self.pointwise("l4_flow_predictor.mlp_convs", [F1[4], emb4])
            qt = self.pose_head("pose_calculator_4", emb4, mask4, None, pose, 3)
            if trace is not None:
                trace.update({"cv3.out": emb, "l4.emb": emb4, "l4.mask": mask4, "l4.qt": qt})
                for l in range(1, 5):
                    trace[f"f1.psa{l}.feats"], trace[f"f2.psa{l}.feats"] = F1[l], F2[l]
                    trace[f"psa{l}.fps_idx"], trace[f"psa{l}.knn_idx"] = lvl_idx[l - 1]

            emb_prev, mask_prev = emb4, mask4
            keep = []      # few-cloud mode: nothing allocated on a branch stream is recycled before the forward ends
            for l in (3, 2, 1):
                p = f"pose_warp_refinement_{l}"
                need(up_ev[l])

                def upconv(which, prev):
                    m = self.set_conv(f"{p}.setupconv_{which}.mlp", X1[l + 1], prev, X1[l], up_idx[l])
                    return m, self.pointwise(f"{p}.setupconv_{which}.post_mlp", [m, F1[l]])

                if side is None:
                    m_f, cf = upconv("features", emb_prev)
                    m_m, cm = upconv("mask", mask_prev)
                    joins = ()
                else:
                    # the two set-upconvs depend on the previous level's outputs only, the searches and the cost volume on
                    # its pose only: three branches of the captured graph instead of an 11-kernel chain
                    ev0 = torch.cuda.Event()
                    ev0.record(main)
                    joins = []
                    outs = []
                    for st, which, prev in ((self._branch[0], "features", emb_prev), (self._branch[1], "mask", mask_prev)):
                        st.wait_event(ev0)
                        with torch.cuda.stream(st):
                            outs.append(upconv(which, prev))
                            ev = torch.cuda.Event()
                            ev.record(st)
                            joins.append(ev)
                    (m_f, cf), (m_m, cm) = outs
                idx_q, warped = self.knn(X2[l], X1[l], 6, warp_qt=qt)           # pose warp fused into the search
                idx_s = self.knn(warped, warped, 4)
                res, e1 = self.cost_volume(f"{p}.cost_volume", warped, F1[l], X2[l], F2[l], idx_q, idx_s)
                if joins:
                    main.wait_event(joins[0])
                ef = self.pointwise(f"{p}.flow_predictor_features.mlp_convs", [F1[l], res, cf])
                if joins:
                    main.wait_event(joins[1])
                em = cm if l == 1 else self.pointwise(f"{p}.flow_predictor_mask.mlp_convs", [cm, ef, F1[l]])
                qt = self.pose_head(f"{p}.pose_calculator", ef, em, qt, pose, l - 1)
                keep += [m_f, cf, m_m, cm, idx_q, warped, idx_s, res, e1, ef, em, qt]
                if trace is not None:
                    trace.update({f"pwr{l}.up_f": cf, f"pwr{l}.up_m": cm, f"pwr{l}.warped": warped, f"pwr{l}.cv": res,
                                  f"pwr{l}.emb": ef, f"pwr{l}.mask": em, f"pwr{l}.qt": qt, f"pwr{l}.idx_q": idx_q,
                                  f"pwr{l}.idx_s": idx_s})
                emb_prev, mask_prev = ef, em
            if side is not None:
                main.wait_stream(side)
        self._sorted = {}
        return pose, mask_prev.permute(0, 2, 1), X1[1]
