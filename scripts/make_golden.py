"""Generate tests/golden/*.npz from the reference's own compiled C++ cores (oracle/_ref, built from
/root/reference by oracle/Makefile). Run in the build container only:

    python scripts/make_golden.py

Inputs: a crop of the reference's demo fragment demo_data/cloud_bin_0.ply (real scan data) and seeded
synthetic clouds. Outputs are stored in REFERENCE order (std::unordered_map iteration order for the
subsampling, nanoflann + std::sort order for the neighbours); the tests canonicalise before comparing.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import native as on  # noqa: E402
from d3feat_b200 import synth  # noqa: E402


def demo_crop(n_raw=30000):
    from utils.ply import read_ply
    d = read_ply("/root/reference/demo_data/cloud_bin_0.ply")
    pts = np.vstack([d["x"], d["y"], d["z"]]).T.astype(np.float32)
    c = pts.mean(0)
    order = np.argsort(np.linalg.norm(pts - c, axis=1), kind="stable")
    return np.ascontiguousarray(pts[np.sort(order[:n_raw])])


def main():
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    on.build(ref=True)
    assert on.have_ref(), "oracle/_ref missing"
    rng = np.random.default_rng(7)

    # ---- 1. grid subsampling on real scan data, two stacked clouds -----------------------------------
    raw = demo_crop()
    a, b = raw[:18000], raw[18000:]
    stacked = np.concatenate([a, b], 0)
    lens = np.array([a.shape[0], b.shape[0]], np.int32)
    sub_p, sub_b = on.ref_batch_subsampling(stacked, lens, 0.05)
    feats = rng.normal(size=(a.shape[0], 5)).astype(np.float32)
    classes = rng.integers(0, 7, size=(a.shape[0], 2)).astype(np.int32)
    wp, wf, wc = on.ref_grid_subsample(a, feats, classes, sampleDl=0.08)
    np.savez_compressed(os.path.join(out, "subsampling_demo.npz"), points=stacked, lengths=lens, dl=np.float32(0.05),
                        sub_points=sub_p, sub_lengths=sub_b, w_points=a, w_features=feats, w_classes=classes,
                        w_dl=np.float32(0.08), w_sub_points=wp, w_sub_features=wf, w_sub_classes=wc)

    # ---- 2. radius neighbours on the subsampled real data (what the pipeline feeds) --------------------
    q = sub_p
    nb = on.ref_batch_neighbors(q, q, sub_b, sub_b, 0.125)
    sub2_p, sub2_b = on.ref_batch_subsampling(q, sub_b, 0.1)
    nb_pool = on.ref_batch_neighbors(sub2_p, q, sub2_b, sub_b, 0.125)
    nb_up = on.ref_batch_neighbors(q, sub2_p, sub_b, sub2_b, 0.25)
    small = q[:600]
    nb_ord = on.ref_ordered_neighbors(small, small, 0.2)
    np.savez_compressed(os.path.join(out, "neighbors_demo.npz"), points=q, lengths=sub_b, radius=np.float32(0.125),
                        neighbors=nb, pool_points=sub2_p, pool_lengths=sub2_b, pool_neighbors=nb_pool,
                        up_neighbors=nb_up, ord_points=small, ord_radius=np.float32(0.2), ord_neighbors=nb_ord)

    # ---- 3. synthetic stacked fragments incl. exact-tie and grid-aligned edge cases --------------------
    frag = np.concatenate([synth.room_fragment(3, 2500), synth.room_fragment(4, 1800)], 0)
    fl = np.array([2500, 1800], np.int32)
    lattice = (np.stack(np.meshgrid(np.arange(12), np.arange(12), np.arange(6), indexing="ij"), -1)
               .reshape(-1, 3).astype(np.float32) * np.float32(0.03))      # exact d2 ties, points on cell faces
    lat_nb = on.ref_batch_neighbors(lattice, lattice, [lattice.shape[0]], [lattice.shape[0]], 0.075)
    lat_sub, lat_b = on.ref_batch_subsampling(lattice, [lattice.shape[0]], 0.06)
    f_nb = on.ref_batch_neighbors(frag, frag, fl, fl, 0.075)
    f_sub, f_b = on.ref_batch_subsampling(frag, fl, 0.06)
    np.savez_compressed(os.path.join(out, "synthetic.npz"), frag=frag, frag_lengths=fl, frag_neighbors=f_nb,
                        frag_sub=f_sub, frag_sub_lengths=f_b, lattice=lattice, lattice_neighbors=lat_nb,
                        lattice_sub=lat_sub, lattice_sub_lengths=lat_b)
    for f in sorted(os.listdir(out)):
        print(f, os.path.getsize(os.path.join(out, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
