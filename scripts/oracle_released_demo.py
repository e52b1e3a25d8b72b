"""Behavioural known-answer test of the TF-graph restatement (oracle/kpconv_np.py) with the RELEASED 3DMatch model:

  reference's demo pair demo_data/cloud_bin_{0,1}.ply  --voxel 0.03-->  pyramid (reference C++ cores, oracle/_ref)
  --> encoder + decoder + detection scores (numpy restatement, weights / BN statistics / kernel points read from
  results/Log_contraloss/snapshots/snap-54 by d3feat_b200/tf_checkpoint.py)  --> keypoints by score, mutual nearest
  neighbours in descriptor space, RANSAC  --> rigid transform.

This is the demo_registration.py flow of the reference without TensorFlow / Open3D. The restatement cannot be compared
with TensorFlow outputs here (TF 1.12 is not installable), but a wrong restatement of ANY block (influence function,
normalisation, BN epsilon, pooling shadow rows, block wiring, variable naming) turns the trained weights into noise:
descriptors stop matching and the two fragments cannot be registered. The script reports the inlier ratio of the
putative matches and the overlap of the aligned clouds. Build container only (needs /root/reference); CPU only.

    python scripts/oracle_released_demo.py [--out tests/golden/released_demo_summary.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
from scipy.spatial import cKDTree

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import native as on                 # noqa: E402
from oracle import kpconv_np as ok              # noqa: E402
from d3feat_b200 import io_utils, tf_checkpoint  # noqa: E402

REF = "/root/reference"


def calibrate(cfg, clouds, keep=0.8):
    """datasets/common.py:572-673 with the reference cores."""
    from d3feat_b200.pyramid import _level_radii
    levels = _level_radii(cfg)
    hist_n = int(np.ceil(4 / 3 * np.pi * (cfg.density_parameter + 1) ** 3))
    hists = np.zeros((len(levels), hist_n), np.int64)
    for c in clouds:
        p, b = c, np.array([c.shape[0]], np.int32)
        for li, lv in enumerate(levels):
            nbm = on.ref_batch_neighbors(p, p, b, b, lv["conv_r"])
            hists[li] += np.bincount(np.sum(nbm < p.shape[0], axis=1), minlength=hist_n)[:hist_n]
            if lv["dl"] is None:
                break
            p, b = on.ref_batch_subsampling(p, b, lv["dl"])
    cs = np.cumsum(hists.T, axis=0)
    return [int(v) for v in np.sum(cs < keep * cs[hist_n - 1, :], axis=0)]


def describe(cfg, params, limits, pts):
    lens = np.array([pts.shape[0]], np.int32)
    inputs = ok.descriptor_input_pyramid(cfg, pts, lens, limits, on.ref_batch_neighbors, on.ref_batch_subsampling)
    inputs["features"] = np.ones((pts.shape[0], 1), np.float32)
    orc = ok.EncoderOracle(cfg, params, np.float32)
    return orc.decoder(inputs, orc.encoder(inputs), return_scores=True)


def kabsch(a, b):
    ca, cb = a.mean(0), b.mean(0)
    u, _, vt = np.linalg.svd((a - ca).T @ (b - cb))
    d = np.sign(np.linalg.det(vt.T @ u.T))
    r = vt.T @ np.diag([1, 1, d]) @ u.T
    return r, cb - r @ ca


def ransac(src, dst, thr=0.05, iters=20000, seed=0):
    rng = np.random.default_rng(seed)
    best = (0, np.eye(3), np.zeros(3))
    n = src.shape[0]
    for _ in range(iters):
        i = rng.choice(n, 3, replace=False)
        r, t = kabsch(src[i], dst[i])
        inl = int(np.sum(np.linalg.norm(src @ r.T + t - dst, axis=1) < thr))
        if inl > best[0]:
            best = (inl, r, t)
    inl = np.linalg.norm(src @ best[1].T + best[2] - dst, axis=1) < thr
    r, t = kabsch(src[inl], dst[inl])
    return r, t, inl


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--keypts", type=int, default=2500)
    args = ap.parse_args()
    on.build(ref=True)
    cfg = io_utils.load_config(os.path.join(REF, "results", "Log_contraloss"))
    params = tf_checkpoint.load_params(os.path.join(REF, "results", "Log_contraloss", "snapshots", "snap-54"))
    clouds = []
    for i in (0, 1):
        raw = io_utils.read_ply_points(os.path.join(REF, "demo_data", "cloud_bin_%d.ply" % i))
        sub, _ = on.ref_batch_subsampling(raw, np.array([raw.shape[0]], np.int32), cfg.first_subsampling_dl)
        clouds.append(sub)
        print("cloud_bin_%d: %d raw -> %d points" % (i, raw.shape[0], sub.shape[0]))
    limits = calibrate(cfg, clouds)
    print("neighbourhood limits:", limits)
    desc, score = [], []
    for c in clouds:
        t0 = time.time()
        d, s = describe(cfg, params, limits, c)
        print("  described %d points in %.1f s; |d| in [%.4f, %.4f]" % (c.shape[0], time.time() - t0,
                                                                       np.linalg.norm(d, axis=1).min(),
                                                                       np.linalg.norm(d, axis=1).max()))
        desc.append(d)
        score.append(s[:, 0])
    kp = [np.argsort(s)[-args.keypts:] for s in score]
    d0, d1 = desc[0][kp[0]], desc[1][kp[1]]
    nn01 = cKDTree(d1).query(d0)[1]
    nn10 = cKDTree(d0).query(d1)[1]
    mutual = np.nonzero(nn10[nn01] == np.arange(d0.shape[0]))[0]
    src, dst = clouds[0][kp[0]][mutual], clouds[1][kp[1]][nn01[mutual]]
    r, t, inl = ransac(src, dst)
    before = float(np.mean(cKDTree(clouds[1]).query(clouds[0])[0] < 0.05))
    after = float(np.mean(cKDTree(clouds[1]).query(clouds[0] @ r.T + t)[0] < 0.05))
    # the same pipeline with the weights shuffled inside every tensor: what a broken restatement looks like
    rng = np.random.default_rng(1)
    broken = {k: (rng.permutation(v.reshape(-1)).reshape(v.shape) if k.endswith("weights") else v)
              for k, v in params.items()}
    db = [describe(cfg, broken, limits, c)[0] for c in clouds]
    b0, b1 = db[0][kp[0]], db[1][kp[1]]
    bn01 = cKDTree(b1).query(b0)[1]
    bmut = np.nonzero(cKDTree(b0).query(b1)[1][bn01] == np.arange(b0.shape[0]))[0]
    bgood = float(np.mean(np.linalg.norm(clouds[0][kp[0]][bmut] @ r.T + t - clouds[1][kp[1]][bn01[bmut]], axis=1) < 0.1)) \
        if bmut.size else 0.0
    out = dict(points=[int(c.shape[0]) for c in clouds], limits=limits, keypoints=args.keypts,
               mutual_matches=int(mutual.size), ransac_inliers=int(inl.sum()),
               inlier_ratio=float(inl.sum() / max(mutual.size, 1)),
               overlap_before=before, overlap_after=after,
               rotation_deg=float(np.degrees(np.arccos(np.clip((np.trace(r) - 1) / 2, -1, 1)))),
               translation_m=float(np.linalg.norm(t)),
               shuffled_weights_mutual_matches=int(bmut.size), shuffled_weights_correct_match_ratio=bgood)
    print(json.dumps(out, indent=1))
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
