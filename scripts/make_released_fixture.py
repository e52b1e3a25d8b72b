"""Build container only: pack the RELEASED 3DMatch model and the reference's demo pair into ONE local fixture that travels
to the GPU box but is never committed (tests/golden_local/ is git-ignored, not gpurun-ignored):

    tests/golden_local/released_3dmatch_full.npz
        params/<variable name>   every tensor of results/Log_contraloss/snapshots/snap-54 (weights, BN statistics,
                                 kernel points) as read by d3feat_b200/tf_checkpoint.py
        cloud0, cloud1           demo_data/cloud_bin_{0,1}.ply voxelised at first_subsampling_dl by the reference's own
                                 grid_subsampling core (oracle/_ref)
        architecture / scalars   results/Log_contraloss/parameters.txt

tests/test_gpu_real_configs.py::test_released_model_registers_the_demo_pair_on_the_gpu runs the whole CUDA path on it.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import native as on                  # noqa: E402
from d3feat_b200 import io_utils, tf_checkpoint  # noqa: E402

REF = "/root/reference"


def main():
    on.build(ref=True)
    log = os.path.join(REF, "results", "Log_contraloss")
    cfg = io_utils.load_config(log)
    params = tf_checkpoint.load_params(os.path.join(log, "snapshots", "snap-54"))
    out = {"params|" + k.replace("/", "|"): np.asarray(v, np.float32) for k, v in params.items()}
    for i in (0, 1):
        raw = io_utils.read_ply_points(os.path.join(REF, "demo_data", "cloud_bin_%d.ply" % i))
        sub, _ = on.ref_batch_subsampling(raw, np.array([raw.shape[0]], np.int32), cfg.first_subsampling_dl)
        out["cloud%d" % i] = sub
    out["architecture"] = np.array(" ".join(cfg.architecture))
    for k in ("first_subsampling_dl", "density_parameter", "KP_extent", "first_features_dim", "num_kernel_points",
              "in_features_dim"):
        out["cfg|" + k] = np.array(getattr(cfg, k))
    for k in ("KP_influence", "convolution_mode", "fixed_kernel_points"):
        out["cfg|" + k] = np.array(getattr(cfg, k))
    d = os.path.join(ROOT, "tests", "golden_local")
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, "released_3dmatch_full.npz")
    np.savez(path, **out)
    print(path, "%.1f MB" % (os.path.getsize(path) / 1e6), len(params), "tensors,", out["cloud0"].shape, out["cloud1"].shape)


if __name__ == "__main__":
    main()
