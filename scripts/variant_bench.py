"""GPU: time the level-0 KPConv layers for the library given by D3F_LIB (kernel-tuning experiments)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from d3feat_b200 import synth, convolution_ops as co
from d3feat_b200.encoder import KPFCNN
dev = torch.device("cuda", 0)
cfg = synth.Config(architecture=synth.ARCH_ENCODER)
params = synth.make_params(cfg, 0)
clouds = [synth.room_fragment(f, 30000) for f in range(8)]
P = np.concatenate(clouds, 0); L = np.array([c.shape[0] for c in clouds], np.int32)
enc = KPFCNN(cfg, params, [40] * 5, device=dev)
inputs = enc.build_inputs(torch.from_numpy(P).to(dev), torch.from_numpy(L).to(dev))
def bench(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
res = []
q0, idx0 = inputs["points"][0], inputs["neighbors"][0]
ones = torch.ones((q0.shape[0], 1), device=dev)
Kp, W = enc.store.get("layer_0/simple_0/kernel_points"), enc.store.get("layer_0/simple_0/weights")
res.append("1->64@%d: %.3f ms" % (q0.shape[0], bench(lambda: co.KPConv_ops(q0, q0, idx0, ones, Kp, W, 0.03, "linear", "sum"))))
for lvl, (Cin, name) in enumerate([(32, "layer_0/resnetb_1/conv2"), (64, "layer_1/resnetb_0/conv2"), (128, "layer_2/resnetb_0/conv2"),
                                   (256, "layer_3/resnetb_0/conv2"), (512, "layer_4/resnetb_0/conv2")]):
    q, idx = inputs["points"][lvl], inputs["neighbors"][lvl]
    feat = torch.randn((q.shape[0], Cin), device=dev)
    Kp, W = enc.store.get(name + "/kernel_points"), enc.store.get(name + "/weights")
    ext = 0.03 * 2 ** lvl
    res.append("%d->%d@%d: %.3f ms" % (Cin, Cin, q.shape[0], bench(lambda: co.KPConv_ops(q, q, idx, feat, Kp, W, ext, "linear", "sum"))))
print(os.environ.get("D3F_LIB", "default").split("/")[-1], " | ".join(res))
