"""Small driver for ncu captures: the dominant kernels at bench shapes (8 x 30k-pt fragments, level 0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from d3feat_b200 import synth, convolution_ops as co
from d3feat_b200.encoder import KPFCNN
dev = torch.device("cuda", 0)
cfg = synth.Config(architecture=synth.ARCH_ENCODER)
params = synth.make_params(cfg, 0)
F = int(os.environ.get("FRAGS", "8"))
clouds = [synth.room_fragment(f, 30000) for f in range(F)]
P = np.concatenate(clouds, 0); L = np.array([c.shape[0] for c in clouds], np.int32)
enc = KPFCNN(cfg, params, [40] * 5, device=dev)
inputs = enc.build_inputs(torch.from_numpy(P).to(dev), torch.from_numpy(L).to(dev))
q, idx = inputs["points"][0], inputs["neighbors"][0]
N = q.shape[0]
feat = torch.randn((N, 32), device=dev)
Kp = enc.store.get("layer_0/resnetb_1/conv2/kernel_points"); W = enc.store.get("layer_0/resnetb_1/conv2/weights")
x32 = torch.randn((N, 32), device=dev); w = enc.store.get("layer_0/resnetb_1/conv3/weights")
x64 = torch.randn((N, 64), device=dev); wsc = enc.store.get("layer_0/resnetb_1/shortcut/weights")
ONLY = os.environ.get("ONLY", "")
torch.cuda.synchronize()
torch.cuda.profiler.start()      # with `ncu --profile-from-start off` only the operator launches are recorded
for it in range(3):
    if ONLY in ("", "kpconv"):
        co.KPConv_ops(q, q, idx, feat, Kp, W, 0.03, "linear", "sum")
    if ONLY in ("", "unary"):
        co.unary_convolution(x32, w)                                  # 240000 x 32 -> 128 (skinny single-stage variant)
        ones, zeros = torch.ones(128, device=dev), torch.zeros(128, device=dev)
        co.unary_pair_convolution(x32, w, (ones, zeros), x64, wsc, (ones, zeros), 0.2)   # conv3 + shortcut, K = 32 + 64
torch.cuda.synchronize()
torch.cuda.profiler.stop()
