"""GPU debug: run the fused KPConv several times on the same inputs and characterise any run-to-run difference."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from d3feat_b200 import convolution_ops as co
from test_gpu_kpconv import make_case
dev = torch.device("cuda", 0)
rng = np.random.default_rng(6040)
q, s, idx, f, Kp, W = make_case(rng, 6000, 6000, 40, 32, 32, extent=0.05)
args = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (q, s, idx, f, Kp, W)]
os.environ["D3F_FUSED_KPCONV"] = "0"
two = co.KPConv_ops(*args, 0.05, "linear", "sum").cpu().numpy()
den = np.abs(two).max()
os.environ["D3F_FUSED_KPCONV"] = "1"
for dbg in ("0", "1"):
  os.environ["D3F_FUSED_DBG"] = dbg
  print("==== D3F_FUSED_DBG =", dbg)
  outs = [co.KPConv_ops(*args, 0.05, "linear", "sum").cpu().numpy() for _ in range(6)]
  for i, o in enumerate(outs):
    d = np.abs(o - outs[0])
    rows = np.nonzero(d.max(1) > 0)[0]
    e2 = np.abs(o - two).max(1) / den
    print("run", i, "rows differing from run 0:", len(rows), rows[:12].tolist(), "max diff/den %.2e" % (d.max() / den),
          "| vs two-kernel: max %.2e, rows > 1e-5: %s" % (e2.max(), np.nonzero(e2 > 1e-5)[0][:12].tolist()))
    if len(rows):
        r = rows[0]
        print("   row", r, "tile", r // 48, "row in tile", r % 48, "cols differing", np.nonzero(d[r] > 0)[0].tolist()[:16])
