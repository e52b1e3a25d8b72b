"""Small-size pass over every kernel of the library, meant to run under compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from d3feat_b200 import synth, tf_custom_ops as ops, cpp_subsampling, convolution_ops as co
from d3feat_b200.encoder import KPFCNN
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
# native ops
P = np.concatenate([synth.room_fragment(0, 1500), synth.room_fragment(1, 1100)], 0)
L = np.array([1500, 1100], np.int32)
tp, tl = torch.from_numpy(P).to(dev), torch.from_numpy(L).to(dev)
nb = ops.batch_ordered_neighbors(tp, tp, tl, tl, 0.1)
sp, sb = ops.batch_grid_subsampling(tp, tl, 0.06)
ops.ordered_neighbors(tp[:300], tp[:300], 0.2)
cpp_subsampling.compute(P[:1500], features=rng.normal(size=(1500, 3)).astype(np.float32), classes=rng.integers(0, 4, (1500,)).astype(np.int32), sampleDl=0.08)
# dense rows -> generic path of the query kernel
D = torch.from_numpy(rng.uniform(0, 0.15, (700, 3)).astype(np.float32)).to(dev)
n700 = torch.tensor([700], dtype=torch.int32, device=dev)
ops.batch_ordered_neighbors(D, D, n700, n700, 0.14, max_cols=32)
# encoder + decoder (every KPConv variant incl. Cin=1, 32, 64, 128, 256, 512; TC GEMM all tile widths; split-K; pools)
cfg = synth.Config()
params = synth.make_params(cfg, 0)
enc = KPFCNN(cfg, params, [30] * 5, device=dev)
out = enc(np.concatenate([synth.room_fragment(2, 2500), synth.room_fragment(3, 2000)], 0), np.array([2500, 2000], np.int32), decoder=True)
# deformable + generic influences / closest mode + CUDA-core fallback paths
cfgd = synth.Config(architecture=synth.ARCH_KITTI_DEFORM, first_subsampling_dl=0.3, first_features_dim=32, modulated=True)
encd = KPFCNN(cfgd, synth.make_params(cfgd, 1), [24] * 5, device=dev)
encd(synth.lidar_scan(0, 3000), np.array([3000], np.int32))
q = torch.from_numpy(rng.uniform(0, 1, (400, 3)).astype(np.float32)).to(dev)
idx = torch.from_numpy(rng.integers(0, 401, (400, 20)).astype(np.int32)).to(dev)
for Cin, Cout in [(32, 32), (48, 40), (3, 8)]:
    f = torch.randn(400, Cin, device=dev); W = torch.randn(15, Cin, Cout, device=dev); Kp = torch.randn(15, 3, device=dev) * 0.1
    for infl in ("constant", "linear", "gaussian"):
        for mode in ("sum", "closest"):
            co.KPConv_ops(q, q, idx, f, Kp, W, 0.12, infl, mode)
# skinny single-stage GEMM variant (needs > 592 output tiles), pair GEMM, all-shadow pooled row, pipeline ring
co.unary_convolution(torch.randn(80000, 32, device=dev), torch.randn(32, 64, device=dev))
one = lambda c: (torch.ones(c, device=dev), torch.zeros(c, device=dev))
co.unary_pair_convolution(torch.randn(5000, 32, device=dev), torch.randn(32, 128, device=dev), one(128),
                          torch.randn(5000, 64, device=dev), torch.randn(64, 128, device=dev), one(128), 0.2)
from d3feat_b200 import network_blocks as nbk
pi = torch.from_numpy(rng.integers(0, 401, (100, 9)).astype(np.int32)).to(dev)
pi[3] = 400
nbk.ind_max_pool(torch.randn(400, 64, device=dev), pi)
from d3feat_b200.encoder import BatchPipeline
cfge = synth.Config(architecture=synth.ARCH_ENCODER)
ence = KPFCNN(cfge, synth.make_params(cfge, 2), [30] * 5, device=dev)
pipe = BatchPipeline(ence)
bp, bl = np.concatenate([synth.room_fragment(4, 1500), synth.room_fragment(5, 1200)], 0), np.array([1500, 1200], np.int32)
pipe.prime(bp, bl)
for _ in range(4):
    pipe.step(bp, bl)
pipe.step(None, None)
pipe.drain()
# ---- round 2: static (sync-free) pyramid + CUDA graph pipeline, device-side row counts, runtime-K KPConv, the fused
# persistent KPConv kernel, the streaming GEMM, the pared stage-1 / first-layer kernels (run by the encoders above)
from d3feat_b200.encoder import GraphPipeline
tbp, tbl = torch.from_numpy(bp).to(dev), torch.from_numpy(bl).to(dev)
gp = GraphPipeline.for_batch(ence, tbp, tbl)
gp.prime(tbp, tbl)
for _ in range(3):
    gp.step(tbp, tbl)
gp.step(None, None)
gp.check()
rows = torch.tensor([350], dtype=torch.int32, device=dev)
co.unary_convolution(torch.randn(400, 64, device=dev), torch.randn(64, 32, device=dev), rows=rows)
for K in (7, 23):
    co.KPConv_ops(q, q, idx, torch.randn(400, 16, device=dev), torch.randn(K, 3, device=dev) * 0.1, torch.randn(K, 16, 24, device=dev), 0.12, "linear", "sum")
os.environ["D3F_FUSED_KPCONV"] = "1"
qf = torch.from_numpy(rng.uniform(0, 1, (4000, 3)).astype(np.float32)).to(dev)   # >= 3552 queries: the fused kernel's floor
idf = torch.from_numpy(rng.integers(0, 4001, (4000, 21)).astype(np.int32)).to(dev)
co.KPConv_ops(qf, qf, idf, torch.randn(4000, 32, device=dev), torch.randn(15, 3, device=dev) * 0.1, torch.randn(15, 32, 32, device=dev), 0.12, "linear", "sum")
os.environ["D3F_FUSED_KPCONV"] = "0"
os.environ["D3F_TC_STREAM"] = "1"
co.unary_convolution(torch.randn(38001, 256, device=dev), torch.randn(256, 48, device=dev))
os.environ["D3F_TC_STREAM"] = "0"
co.USE_TENSOR_CORES = False
co.unary_convolution(torch.randn(300, 36, device=dev), torch.randn(36, 50, device=dev))
co.KPConv_ops(q, q, idx, torch.randn(400, 32, device=dev), torch.randn(15, 3, device=dev) * 0.1, torch.randn(15, 32, 32, device=dev), 0.12, "linear", "sum")
torch.cuda.synchronize()
print("sanitize pass done")
