"""GPU: where does a bench step spend its time? Host-synchronised timing of each phase / op (CUDA events),
plus the un-synchronised end-to-end time, to separate device time from host launch overhead."""
import os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from d3feat_b200 import synth, pyramid, _lib
from d3feat_b200 import convolution_ops as co, network_blocks as nb, tf_custom_ops as ops
from d3feat_b200.encoder import KPFCNN

dev = torch.device("cuda", 0)
F = int(os.environ.get("FRAGS", "8"))
cfg = synth.Config(architecture=synth.ARCH_ENCODER)
params = synth.make_params(cfg, 0)
clouds = [synth.room_fragment(f, 30000) for f in range(F)]
P = np.concatenate(clouds, 0); L = np.array([c.shape[0] for c in clouds], np.int32)
Pd, Ld = torch.from_numpy(P).to(dev), torch.from_numpy(L).to(dev)
bbox = np.concatenate([P.min(0), P.max(0)]).astype(np.float32)
enc = KPFCNN(cfg, params, [40] * 5, device=dev)

def sync_time(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3

def host_time(fn, n=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return (t1 - t0) / n * 1e3

print("full step        : %.3f ms (host-side issue time %.3f ms)" % (sync_time(lambda: enc(Pd, Ld, bbox=bbox, decoder=False)), host_time(lambda: enc(Pd, Ld, bbox=bbox, decoder=False))))
print("pyramid          : %.3f ms (host %.3f)" % (sync_time(lambda: enc.build_inputs(Pd, Ld, bbox=bbox)), host_time(lambda: enc.build_inputs(Pd, Ld, bbox=bbox))))
inputs = enc.build_inputs(Pd, Ld, bbox=bbox)
print("encoder          : %.3f ms (host %.3f)" % (sync_time(lambda: enc.encode(inputs)), host_time(lambda: enc.encode(inputs))))
print("levels:", [int(p.shape[0]) for p in inputs["points"]])

# per-op device time inside the encoder: wrap the ops
acc = collections.OrderedDict()
def wrap(mod, name):
    orig = getattr(mod, name)
    def f(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = orig(*a, **k)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
        if name == "KPConv_ops":
            key = "KPConv Nq=%d Cin=%d Cout=%d" % (a[0].shape[0], a[5].shape[1], a[5].shape[2])
        elif name == "unary_convolution":
            key = "unary N=%d %d->%d" % (a[0].shape[0], a[1].shape[0], a[1].shape[1])
        else:
            key = name + " N=%d C=%d" % (a[1].shape[0], a[0].shape[1])
        acc.setdefault(key, []).append(dt)
        return r
    setattr(mod, name, f)
wrap(co, "KPConv_ops"); wrap(co, "unary_convolution"); wrap(nb, "ind_max_pool")
for _ in range(3):
    enc.encode(inputs)
tot = 0
for k, v in acc.items():
    m = float(np.mean(v[len(v)//3:])); n = len(v) // 3
    tot += m * n
    print("  %-40s x%d  %.3f ms each" % (k, n, m))
print("  sum of synchronised op times: %.3f ms" % tot)

# pyramid pieces
def grid_fill():
    g = ops.NeighborGrid(Pd, Ld, 0.075, bbox)
    return g.fill(Pd, Ld, 40, Pd.shape[0])
print("L0 grid build + fill(40 cols): %.3f ms" % sync_time(grid_fill))
g = ops.NeighborGrid(Pd, Ld, 0.075, bbox)
print("L0 fill only                 : %.3f ms" % sync_time(lambda: g.fill(Pd, Ld, 40, Pd.shape[0])))
print("L0 grid build only           : %.3f ms" % sync_time(lambda: ops.NeighborGrid(Pd, Ld, 0.075, bbox)))
print("L0 subsample dl=0.06         : %.3f ms" % sync_time(lambda: ops.batch_grid_subsampling(Pd, Ld, 0.06, bbox=bbox)))
print("launches per step:", end=" ")
n0 = _lib.launch_count(); enc(Pd, Ld, bbox=bbox, decoder=False); print(_lib.launch_count() - n0)
