"""tests/golden/released_3dmatch_head.npz: the variables of the first four blocks (simple, resnetb, resnetb_strided,
resnetb) of the reference's RELEASED 3DMatch model, read from results/Log_contraloss/snapshots/snap-54 with
d3feat_b200/tf_checkpoint.py. Run in the build container only (the GPU box has no /root/reference):

    python scripts/make_golden_released.py

The fixture lets the GPU parity tests run trained weights, trained BN statistics and the trained run's kernel points
on real scan data (tests/golden/subsampling_demo.npz) instead of seeded random parameters.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from d3feat_b200 import tf_checkpoint as ck  # noqa: E402

SCOPES = ("layer_0/simple_0/", "layer_0/resnetb_1/", "layer_0/resnetb_strided_2/", "layer_1/resnetb_0/")


def main():
    params = ck.load_params("/root/reference/results/Log_contraloss/snapshots/snap-54")
    keep = {k: v for k, v in params.items() if k.startswith(SCOPES)}
    out = os.path.join(ROOT, "tests", "golden", "released_3dmatch_head.npz")
    np.savez_compressed(out, **{k.replace("/", "|"): v for k, v in keep.items()})
    print(out, len(keep), "tensors", os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
