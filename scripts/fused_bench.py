"""GPU: the Cin = 32 KPConv layers at the bench shapes, fused persistent kernel vs the two-kernel path (same library,
D3F_FUSED_KPCONV toggles). Prints one line per (layer, path): median ms over 20 L2-flushed runs, algorithmic GB/s and the
fraction of the measured HBM peak."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from d3feat_b200 import synth, convolution_ops as co
from d3feat_b200.encoder import KPFCNN
import bench as B
dev = torch.device("cuda", 0)
cfg = synth.Config(architecture=synth.ARCH_ENCODER)
params = synth.make_params(cfg, 0)
clouds = [synth.room_fragment(f, 30000) for f in range(int(os.environ.get("FRAGS", "8")))]
P = np.concatenate(clouds, 0); L = np.array([c.shape[0] for c in clouds], np.int32)
enc = KPFCNN(cfg, params, [40] * 5, device=dev)
inputs = enc.build_inputs(torch.from_numpy(P).to(dev), torch.from_numpy(L).to(dev))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
peak, _ = B.peaks()

def med(fn, n=20):
    for _ in range(3): fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        flush.fill_(1); a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev]))

cases = [("L0 32->32 conv", inputs["points"][0], inputs["points"][0], inputs["neighbors"][0], "layer_0/resnetb_1/conv2", 0.03),
         ("L0->1 32->32 strided", inputs["points"][1], inputs["points"][0], inputs["pools"][0], "layer_0/resnetb_strided_2/conv2", 0.03)]
for name, q, s, idx, scope, ext in cases:
    feat = torch.randn((s.shape[0], 32), device=dev)
    Kp, W = enc.store.get(scope + "/kernel_points"), enc.store.get(scope + "/weights")
    ab = B.kpconv_algorithmic_bytes(int(idx.shape[0]), int(idx.shape[1]), 15, 32, int(W.shape[2]))
    outs = {}
    for mode in ("1", "0"):
        os.environ["D3F_FUSED_KPCONV"] = mode
        ms = med(lambda: co.KPConv_ops(q, s, idx, feat, Kp, W, ext, "linear", "sum"))
        outs[mode] = co.KPConv_ops(q, s, idx, feat, Kp, W, ext, "linear", "sum")
        print(json.dumps(dict(layer=name, Nq=int(idx.shape[0]), path="fused" if mode == "1" else "two-kernel", ms=ms,
                              gbs=ab / ms / 1e6, frac=ab / ms / 1e6 / peak)))
    d = (outs["1"] - outs["0"]).abs().max().item() / outs["0"].abs().max().item()
    print(json.dumps(dict(layer=name, fused_vs_two_kernel_rel=d)))
