"""BASELINE.json configs[2]: one KITTI-shaped 120k-point scan through the deformable architecture (not the bench.py
headline; a side measurement for DESIGN.md). Prints one JSON line: ms per scan, points/s, per-stage split."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from d3feat_b200 import synth                     # noqa: E402
from d3feat_b200.encoder import KPFCNN            # noqa: E402

LIMITS = [40, 40, 40, 60, 40]


def main():
    dev = torch.device("cuda", 0)
    cfg = synth.Config(architecture=synth.ARCH_KITTI_DEFORM, first_subsampling_dl=0.04, first_features_dim=32)
    cloud = synth.lidar_scan(1, 120000, dl=0.04)
    params = synth.make_params(cfg, 1)
    enc = KPFCNN(cfg, params, LIMITS, device=dev)
    P = torch.from_numpy(cloud).to(dev)
    L = torch.tensor([cloud.shape[0]], dtype=torch.int32, device=dev)
    bbox = np.concatenate([cloud.min(0), cloud.max(0)]).astype(np.float32)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn, steps=10, warmup=3):
        for _ in range(warmup):
            fn()
        ts = []
        for _ in range(steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.mean(ts))

    inputs = enc.build_inputs(P, L, bbox=bbox)
    t_all = timed(lambda: enc(P, L, bbox=bbox, decoder=False))
    t_pyr = timed(lambda: enc.build_inputs(P, L, bbox=bbox))
    t_enc = timed(lambda: enc.encode(inputs))
    sizes = [int(p.shape[0]) for p in inputs["points"]]
    print(json.dumps(dict(workload="KITTI-shaped 120k-pt scan, deformable KPConv (last 3 blocks), 1 GPU",
                          ms_per_scan=t_all, points_per_s=cloud.shape[0] / t_all * 1e3, pyramid_ms=t_pyr,
                          encoder_ms=t_enc, level_sizes=sizes)))


if __name__ == "__main__":
    main()
