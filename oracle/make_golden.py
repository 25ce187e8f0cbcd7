"""oracle/make_golden.py -- generates tests/golden/*.npz by running the UNMODIFIED reference
(PW/pwclo_net.py via oracle/ref_shim.py) on CPU in the build container.  Re-run with
`python -m oracle.make_golden`; needs /root/reference.

What is pinned (all produced by reference code, none by our port):
  * pose_params [B,4,7] of the full forward (PW/pwclo_net.py:205-207) and the log dict;
  * every FPS index tensor (9 calls, through the `_ext.furthest_point_sampling` call sites,
    P2/pointnet2_modules.py:200-206);
  * for every knn_point call (23, P2/pytorch_utils.py:32-49): the per-row index sums, the sorted
    rows of a strided row subset;
  * strided slices of the module outputs of psa_1..4 (both frames), cost_volume,
    flow_feature_encoding, the three PoseWarpRefinement modules.
Inputs and weights are NOT stored: they are regenerated from seeds by
pwclonet_pylidarslam_b200/synthetic.py; their SHA-256 is stored so that generator drift is detected.
"""
import hashlib
import os

import numpy as np
import torch

from oracle import cpu_ops, ref_shim
from pwclonet_pylidarslam_b200 import synthetic as syn

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


class RecordingExt(type(cpu_ops.torch_ext)):
    def __init__(self):
        self.fps = []

    def furthest_point_sampling(self, points, nsamples):
        out = super().furthest_point_sampling(points, nsamples)
        self.fps.append(out.numpy().copy())
        return out


def run(first_pair, n_pairs, n_points, weight_seed, tag):
    ext = RecordingExt()
    net = ref_shim.load_reference(ext)
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    w = syn.make_state_dict(shapes, seed=weight_seed)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    x1, x2, gt = syn.make_batch(first_pair, n_pairs, n_points)

    mods = ref_shim.reference_modules()
    ptu = mods["pytorch_utils"]
    knn_calls = []
    orig_knn = ptu.knn_point

    def rec_knn(nsample, xyz, new_xyz):
        a, b = orig_knn(nsample, xyz, new_xyz)
        knn_calls.append(b.numpy().copy())
        return a, b

    outs = {}

    def hook(name):
        def f(mod, inp, out):
            outs.setdefault(name, []).append(out)
        return f

    hs = []
    for name in ["psa_1", "psa_2", "psa_3", "psa_4", "cost_volume", "flow_feature_encoding",
                 "pose_warp_refinement_3", "pose_warp_refinement_2", "pose_warp_refinement_1"]:
        hs.append(getattr(net, name).register_forward_hook(hook(name)))
    ptu.knn_point = rec_knn
    try:
        with torch.no_grad():
            pose, log = net(torch.from_numpy(x1), None, torch.from_numpy(x2), None)
    finally:
        ptu.knn_point = orig_knn
        for h in hs:
            h.remove()

    g = {"pose": pose.numpy(), "log_embedding_mask": log["embedding_mask"].numpy()[:, ::16],
         "input_sha": np.array(sha(x1, x2)), "weights_sha": np.array(sha(*[w[k] for k in sorted(w)])),
         "meta": np.array([first_pair, n_pairs, n_points, weight_seed])}
    assert len(ext.fps) == 9 and len(knn_calls) == 23, (len(ext.fps), len(knn_calls))
    for i, f in enumerate(ext.fps):
        g[f"fps_{i}"] = f.astype(np.int16 if f.max() < 32768 else np.int32)
    for i, k in enumerate(knn_calls):
        g[f"knn_{i}_rowsum"] = k.sum(-1).astype(np.int32)
        stride = max(1, k.shape[1] // 32)
        g[f"knn_{i}_rows"] = np.sort(k[:, ::stride], axis=-1).astype(np.int16 if k.max() < 32768 else np.int32)
    for l in range(1, 5):
        for fr in range(2):
            new_xyz, feats = outs[f"psa_{l}"][fr]
            g[f"psa{l}_f{fr + 1}_feats"] = feats.numpy()[:, :, ::max(1, feats.shape[2] // 64)]
    g["cv3_out"] = outs["cost_volume"][0].numpy()[:, :, ::4]
    g["ffe_feats"] = outs["flow_feature_encoding"][0][1].numpy()
    for l in (3, 2, 1):
        q, t, ef, em = outs[f"pose_warp_refinement_{l}"][0]
        g[f"pwr{l}_q"] = q.numpy()
        g[f"pwr{l}_t"] = t.numpy()
        g[f"pwr{l}_emb"] = ef.numpy()[:, :, ::max(1, ef.shape[2] // 64)]
        g[f"pwr{l}_mask"] = em.numpy()[:, :, ::max(1, em.shape[2] // 64)]
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, f"{tag}.npz")
    np.savez_compressed(path, **g)
    print(path, os.path.getsize(path) // 1024, "KiB", "pose[0,0] =", pose[0, 0].numpy())


def ops_golden():
    """Small known-answer vectors for the operators, produced by the C oracle AFTER it was checked
    bit-exact against the reference's CUDA kernels on a B200 (tests/test_ops_gpu.py, oracle/_ref)."""
    rng = np.random.default_rng(2024)
    x = (rng.standard_normal((2, 600, 3)) * [8, 1, 8]).astype(np.float32)
    x[1, 300:] = x[1, :300]          # exact ties
    x[0, ::11] *= 1e-3               # origin-ball points
    g = {"xyz": x, "fps_150": cpu_ops.fps(x, 150), "fps_150_cap1024_noskip": cpu_ops.fps(x, 150, False, 1024)}
    q = x[:, :77] + 0.01
    g["knn8_order0"] = cpu_ops.knn(x, q, 8, 0)
    g["knn8_order1"] = cpu_ops.knn(x, q, 8, 1)
    g["ball_r1_8"] = cpu_ops.ball_query(q, x, 1.0, 8)
    d2, i3 = cpu_ops.three_nn(q, x)
    g["three_nn_d2"], g["three_nn_idx"] = d2, i3
    np.savez_compressed(os.path.join(GOLD, "ops_kat.npz"), **g)


if __name__ == "__main__":
    run(first_pair=0, n_pairs=2, n_points=8192, weight_seed=1, tag="forward_b2_n8192_w1")
    run(first_pair=2, n_pairs=1, n_points=8192, weight_seed=2, tag="forward_b1_n8192_w2")
    ops_golden()
