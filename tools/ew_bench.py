"""Micro-benchmark of the InstanceNorm elementwise kernels on the residual-block shape
(64x64x256, N images): achieved GB/s vs the measured HBM copy bandwidth."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tf2_cyclegan_b200.ops import native as nv

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=12)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--no_flush", action="store_true", help="keep L2 warm between iterations (in-step behaviour)")
ap.add_argument("--direct", action="store_true", help="direct (register-landing) kernels instead of the bulk-copy ring")
a = ap.parse_args()
C = nv.load()
C.set_ew_stream(not a.direct)
n, h, c = a.images, 64, 256
hbm = 6572.9
try:
  hbm = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
  pass
x = nv.Buf(n, h, h, c); x.t.copy_(torch.randn(n, h, h, c, device="cuda").to(torch.bfloat16))
skip = nv.Buf(n, h, h, c, halo=1); skip.t.normal_()
dst = nv.Buf(n, h, h, c, halo=1)
g1 = nv.Buf(n, h, h, c, halo=1); g1.t.normal_()
g2 = nv.Buf(n, h, h, c); g2.t.normal_()
dy = nv.Buf(n, h, h, c)
dx = nv.Buf(n, h, h, c, halo=1)
stats = torch.stack([x.t.float().sum((1, 2)), (x.t.float() ** 2).sum((1, 2))], -1).contiguous()
bst = torch.zeros(n, c, 2, device="cuda")
gam = torch.randn(c, device="cuda"); bet = torch.randn(c, device="cuda")
dg = torch.zeros(c, device="cuda"); db = torch.zeros(c, device="cuda")
inv = 1.0 / (h * h)
E = n * h * h * c * 2
plans = {}
p = C.Plan(); p.add_in_apply("a", x.view(), stats.data_ptr(), gam.data_ptr(), bet.data_ptr(), inv, 1e-3, 1, nv.NULL_VIEW, dst.view(), 1, n, h, h, c); plans["in_apply(relu,reflect)"] = (p, 2 * E)
p = C.Plan(); p.add_in_apply("b", x.view(), stats.data_ptr(), gam.data_ptr(), bet.data_ptr(), inv, 1e-3, 0, skip.view(), dst.view(), 1, n, h, h, c); plans["in_apply(+skip)"] = (p, 3 * E)
p = C.Plan(); p.add_in_bwd_reduce("r", g1.view(), 1, nv.NULL_VIEW, x.view(), stats.data_ptr(), gam.data_ptr(), bet.data_ptr(), inv, 1e-3, 1, dy.view(), bst.data_ptr(), n, h, h, c); plans["in_bwd_reduce(relu,fold)"] = (p, 3 * E)
p = C.Plan(); p.add_in_bwd_reduce("r2", g1.view(), 1, g2.view(), x.view(), stats.data_ptr(), gam.data_ptr(), bet.data_ptr(), inv, 1e-3, 0, dy.view(), bst.data_ptr(), n, h, h, c); plans["in_bwd_reduce(+g2)"] = (p, 4 * E)
p = C.Plan(); p.add_in_bwd_apply("ap", dy.view(), x.view(), stats.data_ptr(), bst.data_ptr(), gam.data_ptr(), inv, 1e-3, dx.view(), dg.data_ptr(), db.data_ptr(), n, h, h, c); plans["in_bwd_apply"] = (p, 3 * E)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
flush2 = torch.zeros(512 << 20, dtype=torch.uint8, device="cuda")   # read after the write: leaves CLEAN lines in L2
s = torch.cuda.current_stream().cuda_stream
print(f"l2_flush={not a.no_flush} kernels={'direct' if a.direct else 'stream'}")
for name, (plan, nbytes) in plans.items():
  for _ in range(3): plan.run(0, -1, s, False)
  ts = []
  for _ in range(a.iters):
    if not a.no_flush:
      flush.zero_(); flush2.sum()   # dirty then clean: the timed kernel does not pay the write-back
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.run(0, -1, s, False); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
  ts.sort(); ms = ts[len(ts) // 2]
  gbs = nbytes / ms / 1e6
  print(f"{name:28s} {ms * 1e3:7.1f} us  {gbs:7.0f} GB/s  {100 * gbs / hbm:5.1f}% of measured HBM copy ({hbm:.0f})")
