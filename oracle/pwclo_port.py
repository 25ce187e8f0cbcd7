"""oracle/pwclo_port.py -- TEST INFRASTRUCTURE ONLY (checker + CPU baseline; never the product path).

A compact CPU restatement ("port") of the reference's PWCLO-Net inference forward, written from
its behaviour and driven by a state dict in the reference's layout.  It exists
because /root/reference cannot travel to the GPU box; it is validated in the build container
against the UNMODIFIED reference (oracle/ref_shim.py) and against the committed golden vectors
(tests/golden, tests/test_oracle_cpu.py).

Every function cites what it restates (paths relative to /root/reference,
P2 = slam/models/Pointnet2_PyTorch/pointnet2_ops_lib/pointnet2_ops, PW = slam/models/PWCLONet).
Sampling / grouping / kNN go through oracle/pointnet2_cpu.c; 1x1 convolutions, batch-norm (eval),
softmax are torch CPU fp32 ops, exactly the ones the reference modules call.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import cpu_ops

BN_EPS = 1e-5  # nn.BatchNorm2d default, P2/pytorch_utils.py:102-105


class Port:
    def __init__(self, state_dict, knn_impl="c", sum_order=0, prefix=""):
        """state_dict: name -> torch tensor / numpy array (reference names, optional `pwclonet.` prefix
        already stripped).  knn_impl 'c' = oracle C kNN ((distance, index) tie rule), 'torch' = the
        reference's materialising formulation (P2/pytorch_utils.py:12-49) for baseline timing."""
        self.sd = {k[len(prefix):] if prefix and k.startswith(prefix) else k:
                   (torch.from_numpy(np.asarray(v)) if not torch.is_tensor(v) else v.detach().cpu())
                   for k, v in state_dict.items()}
        self.knn_impl = knn_impl
        self.sum_order = sum_order
        self.trace = {}
        self.knn_log = []   # every knn index tensor, in call order (23 per forward)
        self.fps_log = []   # every FPS index tensor, in call order (9 per forward)

    # ------------------------------------------------------------------ primitives
    def shared_mlp(self, name, x):
        """P2/pytorch_utils.py:52-83,114-167: [Conv2d 1x1 (no bias) -> BatchNorm2d (eval) -> ReLU]*"""
        i = 0
        while f"{name}.layer{i}.conv.weight" in self.sd:
            p = f"{name}.layer{i}"
            x = F.conv2d(x, self.sd[p + ".conv.weight"])
            x = F.batch_norm(x, self.sd[p + ".bn.bn.running_mean"], self.sd[p + ".bn.bn.running_var"],
                             self.sd[p + ".bn.bn.weight"], self.sd[p + ".bn.bn.bias"], False, 0.0, BN_EPS)
            x = F.relu(x)
            i += 1
        return x

    def conv1d(self, name, x):
        """P2/pytorch_utils.py:170-203 with bn=False, activation=None: nn.Conv1d(k=1) + bias"""
        return F.conv1d(x, self.sd[name + ".conv.weight"], self.sd[name + ".conv.bias"])

    def knn(self, k, xyz, new_xyz):
        """P2/pytorch_utils.py:32-49 -> int32 [B,S,k]"""
        if self.knn_impl == "torch":
            N, S = xyz.shape[1], new_xyz.shape[1]
            diff = new_xyz.unsqueeze(2).repeat(1, 1, N, 1) - xyz.unsqueeze(1).repeat(1, S, 1, 1)
            dist = torch.sqrt(torch.sum(diff ** 2, dim=-1) + 1e-8)
            idx = torch.topk(dist, k, largest=False, dim=-1)[1].int().contiguous()
        else:
            idx = torch.from_numpy(cpu_ops.knn(xyz.numpy(), new_xyz.numpy(), k, self.sum_order))
        self.knn_log.append(idx)
        return idx

    @staticmethod
    def group(feats, idx):
        """P2/pointnet2_utils.py:194-240 -> [B,C,S,K]"""
        return torch.from_numpy(cpu_ops.group_points(feats.contiguous().numpy(), idx.numpy()))

    def fps_gather(self, xyz, npoint):
        """P2/pointnet2_modules.py:198-206: new_xyz = gather(xyz^T, fps(xyz, npoint))^T"""
        idx = cpu_ops.fps(xyz.numpy(), npoint)
        self.fps_log.append(torch.from_numpy(idx))
        flipped = xyz.transpose(1, 2).contiguous()
        new = torch.from_numpy(cpu_ops.gather_points(flipped.numpy(), idx)).transpose(1, 2).contiguous()
        return new, torch.from_numpy(idx)

    # ------------------------------------------------------------------ layers
    def set_conv(self, name, xyz, feats, npoint, nsample, tag=None):
        """PointnetSAModulePWCLONet.forward, P2/pointnet2_modules.py:179-245.
        xyz [B,N,3], feats [B,C,N] or None -> new_xyz [B,S,3], new_feats [B,C',S]"""
        new_xyz, fidx = self.fps_gather(xyz, npoint)
        idx = self.knn(nsample, xyz, new_xyz)
        flipped = xyz.transpose(1, 2).contiguous()
        g_xyz = self.group(flipped, idx)
        diff = g_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if feats is not None:
            x = torch.cat((diff, self.group(feats, idx)), dim=1)       # (xyz_diff, features)  :222
        else:
            x = torch.cat((diff, g_xyz), dim=1)                         # (xyz_diff, grouped_xyz) :233
        x = self.shared_mlp(name + ".mlp_module", x)
        out = x.max(dim=3)[0]
        if tag:
            self.trace[tag + ".fps_idx"] = fidx
            self.trace[tag + ".knn_idx"] = idx
            self.trace[tag + ".new_xyz"] = new_xyz
            self.trace[tag + ".feats"] = out
        return new_xyz, out

    def set_upconv(self, name, xyz2, xyz1, feats2, feats1, nsample=8, tag=None):
        """PointnetFPModulePWCLONet.forward (knn=True), P2/pointnet2_modules.py:459-515.
        xyz2 [B,S,3] dense, xyz1 [B,Sp,3] coarse, feats2 [B,C,S], feats1 [B,C1,Sp] -> [B,64,S]"""
        idx = self.knn(nsample, xyz1, xyz2)
        g_f = self.group(feats1, idx)
        g_xyz = self.group(xyz1.transpose(1, 2).contiguous(), idx)
        diff = g_xyz - xyz2.transpose(1, 2).unsqueeze(-1)
        x = torch.cat((g_f, diff), dim=1)                               # (features, xyz_diff) :491
        x = self.shared_mlp(name + ".mlp", x).max(dim=3)[0]
        x = torch.cat([x, feats2], dim=1).unsqueeze(-1)
        out = self.shared_mlp(name + ".post_mlp", x).squeeze(-1)
        if tag:
            self.trace[tag + ".knn_idx"] = idx
            self.trace[tag + ".out"] = out
        return out

    def cost_volume(self, name, warped_xyz, f1, xyz2, f2, nsample, nsample_q, tag=None):
        """CostVolume.forward, PW/costvolume.py:63-190.  warped_xyz [B,3,S], f1 [B,C,S], xyz2 [B,3,N], f2 [B,C,N]"""
        wt = warped_xyz.permute(0, 2, 1).contiguous()
        x2t = xyz2.permute(0, 2, 1).contiguous()
        idx_q = self.knn(nsample_q, x2t, wt)
        q_xyz = self.group(xyz2, idx_q)
        q_f = self.group(f2, idx_q)
        p_xyz = warped_xyz.unsqueeze(3).expand(-1, -1, -1, nsample_q)
        p_f = f1.unsqueeze(3).expand(-1, -1, -1, nsample_q)
        d = q_xyz - p_xyz
        euc = torch.sqrt(torch.sum(torch.square(d), dim=1, keepdim=True) + 1e-20)
        geo = torch.cat((p_xyz, q_xyz, d, euc), dim=1)                  # 10 channels :105
        x = torch.cat((geo, p_f, q_f), dim=1)                           # :108-110
        x = self.shared_mlp(name + ".mlp_convs", x)
        enc = self.shared_mlp(name + ".mlp_conv_xyz_1", geo)
        wq = F.softmax(self.shared_mlp(name + ".mlp2_convs", torch.cat((enc, x), dim=1)), dim=3)
        e1 = torch.sum(wq * x, dim=3)                                    # first attentive embedding [B,64,S]
        idx = self.knn(nsample, wt, wt)
        c_xyz = self.group(warped_xyz, idx)
        c_e = self.group(e1, idx)
        n_xyz = warped_xyz.unsqueeze(3).expand(-1, -1, -1, nsample)
        n_f = f1.unsqueeze(3).expand(-1, -1, -1, nsample)
        d2 = c_xyz - n_xyz
        euc2 = torch.sqrt(torch.sum(torch.square(d2), dim=1, keepdim=True) + 1e-20)
        geo2 = torch.cat((n_xyz, c_xyz, d2, euc2), dim=1)
        enc2 = self.shared_mlp(name + ".mlp_conv_xyz_2", geo2)
        wp = F.softmax(self.shared_mlp(name + ".mlp3_convs", torch.cat((enc2, n_f, c_e), dim=1)), dim=3)
        out = torch.sum(wp * c_e, dim=3)
        if tag:
            self.trace[tag + ".idx_q"] = idx_q
            self.trace[tag + ".idx_self"] = idx
            self.trace[tag + ".stage1"] = e1
            self.trace[tag + ".out"] = out
        return out

    def flow_predictor(self, name, a, b, c=None):
        """FlowPredictor.forward, PW/flowpredictor.py:53-84: cat along C -> shared MLP per point"""
        x = torch.cat((a, b) if c is None else (a, b, c), dim=1).unsqueeze(3)
        return self.shared_mlp(name + ".mlp_convs", x).squeeze(3)

    def pose_calculator(self, name, feats, mask):
        """PoseCalculator.forward (eval: dropout is identity), PW/pose_calculator.py:47-87 -> q [B,4,1], t [B,3,1]"""
        s = torch.sum(feats * mask, dim=2, keepdim=True)
        big = self.conv1d(name + ".conv1d_q_t", s)
        q = self.conv1d(name + ".conv1d_q", big)
        q = q / (torch.sqrt(torch.sum(q * q, dim=1, keepdim=True) + 1e-10) + 1e-10)
        t = self.conv1d(name + ".conv1d_t", big)
        return q, t

    # ------------------------------------------------------------------ quaternion helpers (scalar first)
    @staticmethod
    def inv_q(q):
        """PW/PWCLO_utils.py:31-39"""
        q2 = torch.sum(q * q, dim=-1, keepdim=True) + 1e-10
        return q * torch.tensor([1, -1, -1, -1]) / q2

    @staticmethod
    def _mul(a, b):
        """Hamilton product on [B,N,4] x [B,1,4] (or reverse), term order of PW/PWCLO_utils.py:82-91,116-126"""
        a0, a1, a2, a3 = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
        b0, b1, b2, b3 = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
        return torch.stack((a0 * b0 - a1 * b1 - a2 * b2 - a3 * b3,
                            a0 * b1 + a1 * b0 + a2 * b3 - a3 * b2,
                            a0 * b2 - a1 * b3 + a2 * b0 + a3 * b1,
                            a0 * b3 + a1 * b2 - a2 * b1 + a3 * b0), dim=-1)

    def warp(self, xyz, q, t):
        """PW/PWCLO_utils.py:42-63: xyz [B,3,N], q [B,4,1], t [B,3,1] -> [B,3,N]"""
        B, _, N = xyz.shape
        qv = q.reshape(B, 1, 4)
        qi = self.inv_q(q.reshape(B, 4)).reshape(B, 1, 4)
        p = torch.cat((torch.zeros(B, 1, N), xyz), dim=1).permute(0, 2, 1)
        r = self._mul(qv, p)          # mul_q_point(q, xyz_)
        s = self._mul(r, qi)          # mul_point_q(., q_inv)
        return s.permute(0, 2, 1)[:, 1:, :] + t

    def pose_warp_refinement(self, name, xyz1, f1, xyz2, f2, xyz1_prev, f1_prev, mask_prev, q_prev, t_prev, last,
                             tag=None):
        """PoseWarpRefinement.forward, PW/pose_warp_refinement.py:82-158 (all xyz [B,3,*])"""
        B = xyz1.shape[0]
        qc = q_prev.reshape(B, 4, 1)
        tc = t_prev.reshape(B, 3, 1)
        x1t = xyz1.permute(0, 2, 1).contiguous()
        xpt = xyz1_prev.permute(0, 2, 1).contiguous()
        cf = self.set_upconv(name + ".setupconv_features", x1t, xpt, f1, f1_prev, tag=tag and tag + ".up_f")
        cm = self.set_upconv(name + ".setupconv_mask", x1t, xpt, f1, mask_prev, tag=tag and tag + ".up_m")
        warped = self.warp(xyz1, qc, tc)
        res = self.cost_volume(name + ".cost_volume", warped, f1, xyz2, f2, 4, 6, tag=tag and tag + ".cv")
        ef = self.flow_predictor(name + ".flow_predictor_features", f1, res, cf)
        em = cm if last else self.flow_predictor(name + ".flow_predictor_mask", cm, ef, f1)
        w = F.softmax(em, dim=2)
        qd, td = self.pose_calculator(name + ".pose_calculator", ef, w)
        q = self._mul(qd.reshape(B, 1, 4), qc.reshape(B, 1, 4)).reshape(B, 4)   # mul_point_q(q_det, q_coarse) :139
        t = self.warp(tc, qd, td).squeeze(2)                                     # :148
        if tag:
            self.trace[tag + ".warped"] = warped
            self.trace[tag + ".emb"] = ef
            self.trace[tag + ".mask"] = em
            self.trace[tag + ".q"] = q
            self.trace[tag + ".t"] = t
        return q, t, ef, em

    # ------------------------------------------------------------------ whole network
    def forward(self, xyz_f1, xyz_f2):
        """PWCLONet.forward, PW/pwclo_net.py:109-207 for xyz-only input. [B,3,N] x2 -> pose_params [B,4,7]"""
        self.trace, self.knn_log, self.fps_log = {}, [], []
        xyz_f1 = torch.as_tensor(xyz_f1)
        xyz_f2 = torch.as_tensor(xyz_f2)
        lv = [(2048, 32), (1024, 32), (256, 16), (64, 16)]
        xs, fs = [[], []], [[], []]
        for fr, x in enumerate((xyz_f1, xyz_f2)):
            xyz = x.permute(0, 2, 1).contiguous()
            feats = None
            for l, (npoint, k) in enumerate(lv):
                xyz, feats = self.set_conv(f"psa_{l + 1}", xyz, feats, npoint, k, tag=f"f{fr + 1}.psa{l + 1}")
                xs[fr].append(xyz)
                fs[fr].append(feats)
        X1 = [x.permute(0, 2, 1).contiguous() for x in xs[0]]   # [B,3,S] per level (index 0 = level 1)
        X2 = [x.permute(0, 2, 1).contiguous() for x in xs[1]]
        emb = self.cost_volume("cost_volume", X1[2], fs[0][2], X2[2], fs[1][2], 4, 32, tag="cv3")
        x14_t, emb4 = self.set_conv("flow_feature_encoding", xs[0][2], emb, 64, 16, tag="ffe")
        x14 = x14_t.permute(0, 2, 1).contiguous()
        mask4 = self.flow_predictor("l4_flow_predictor", fs[0][3], emb4)
        q4, t4 = self.pose_calculator("pose_calculator_4", emb4, F.softmax(mask4, dim=2))
        q4, t4 = q4.squeeze(2), t4.squeeze(2)
        self.trace.update({"l4.mask": mask4, "l4.emb": emb4, "l4.q": q4, "l4.t": t4})
        q3, t3, e3, m3 = self.pose_warp_refinement("pose_warp_refinement_3", X1[2], fs[0][2], X2[2], fs[1][2], x14,
                                                   emb4, mask4, q4, t4, False, tag="pwr3")
        q2, t2, e2, m2 = self.pose_warp_refinement("pose_warp_refinement_2", X1[1], fs[0][1], X2[1], fs[1][1], X1[2],
                                                   e3, m3, q3, t3, False, tag="pwr2")
        q1, t1, e1, m1 = self.pose_warp_refinement("pose_warp_refinement_1", X1[0], fs[0][0], X2[0], fs[1][0], X1[1],
                                                   e2, m2, q2, t2, True, tag="pwr1")
        rows = []
        for q, t in ((q1, t1), (q2, t2), (q3, t3), (q4, t4)):
            qn = q / (torch.sqrt(torch.sum(q * q, dim=-1, keepdim=True) + 1e-10) + 1e-10)
            rows.append(torch.cat((t, qn), dim=-1).reshape(-1, 1, 7))
        pose = torch.cat(rows, dim=1)
        log = {"embedding_mask": torch.linalg.norm(F.softmax(m1, dim=2).permute(0, 2, 1), dim=-1, ord=2),
               "point_cloud": xs[0][0]}
        return pose, log
