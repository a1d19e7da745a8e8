"""Micro-benchmark of the hot GEMM kernels on the residual-block shape
(3x3, 256->256 channels, 64x64, reflect pad): forward conv, flat dgrad (MN-major B),
split-K wgrad. Device-timed with CUDA events, L2 flushed between iterations,
reports TFLOP/s and the fraction of the measured cuBLAS bf16 peak (MEASURED_PEAKS.json).

    python tools/kernel_bench.py [--images 12] [--iters 20] [--only fwd|dgrad|wgrad]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tf2_cyclegan_b200.engine import native_ops as no
from tf2_cyclegan_b200.ops import native as nv

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=12)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--only", default="")
ap.add_argument("--no_flush", action="store_true")
ap.add_argument("--no_stats", action="store_true")
ap.add_argument("--mx_profile", action="store_true", help="stall attribution of the tap-shared MXFP8 kernel (CTA 0)")
ap.add_argument("--layer", default="res", choices=["res", "stem"],
                help="res: 3x3 256->256 @64x64 (default); stem: 7x7 3->64 @256x256 forward (narrow N = 64 kernel)")
a = ap.parse_args()
C = nv.load()
n, h, c = a.images, 64, 256
peaks = {}
try:
  peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
except Exception:
  pass
peak = peaks.get("bf16_tflops", 1590.0)
torch.manual_seed(0)
x = nv.Buf(n, h, h, c, halo=1)
x.t.copy_(torch.randn_like(x.t.float()).to(torch.bfloat16))
w2 = (torch.randn(c, 9 * c, device="cuda") * 0.05).to(torch.bfloat16)
out = nv.Buf(n, h, h, c)
stats = torch.zeros(n, c, 2, device="cuda")
dy = nv.Buf(n, h, h, c, halo=1)
dy.interior().copy_(torch.randn(n, h, h, c, device="cuda").to(torch.bfloat16))
dxp = nv.Buf(n, h, h, c, halo=1)
dw = torch.zeros(c, 9 * c, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
flush2 = torch.zeros(512 << 20, dtype=torch.uint8, device="cuda")   # read after the write: leaves CLEAN lines in L2

plans = {}
if a.layer == "stem":
  # the generator stem: 8-channel pixels, one GEMM tap per kernel row over an overlapping 8-pixel window
  hs = 256
  xin = nv.Buf(n, hs, hs, 8, halo=3, slack_rows=1)
  xin.t.copy_(torch.randn_like(xin.t.float()).to(torch.bfloat16))
  w2s = (torch.randn(64, 448, device="cuda") * 0.05).to(torch.bfloat16)
  bw, bh = nv.tile_shape(hs, hs)
  amap = nv.tmap(xin.ptr(padded=True), [64, hs, xin.hp, n], [8, xin.sh, xin.sn], [64, bw, bh, 1])
  outs = nv.Buf(n, hs, hs, 64)
  stats_s = torch.zeros(n, 64, 2, device="cuda")
  p = C.Plan()
  no.add_conv(p, "stem", [amap], [(0, 0, 0, r, r * 64) for r in range(7)], 64, w2s, outs.view(), hs, hs, n, 64,
              stats=0 if a.no_stats else stats_s.data_ptr())
  plans["stem_fwd"] = (p, 2.0 * n * hs * hs * 49 * 3 * 64)
  a.only = a.only or "stem_fwd"
from tf2_cyclegan_b200.config import ENGINE
ts = ENGINE.tap_share   # tap-shared halo tiles (csrc/conv_gemm_ts.cu)
bw, bh = (8, 16) if ts else nv.tile_shape(h, h)
p = C.Plan()
no.add_conv(p, "fwd", [x.tmap_padded(bw, bh)], no.taps_conv(3, c), c, w2, out.view(), h, h, n, c,
            stats=0 if a.no_stats else stats.data_ptr(), tile=(bw, bh),
            halo_map=(lambda w_, h_: x.tmap_padded(w_, h_)) if ts else None)
plans["fwd"] = (p, 2.0 * n * h * h * 9 * c * c)
# the same conv with MXFP8 operands (csrc/conv_gemm_mx.cu): random E4M3 bytes / moderate scales
x8 = nv.Buf8(n, h, h, c, halo=1)
x8.t.copy_(torch.randint(0, 0x70, x8.t.shape, device="cuda", dtype=torch.uint8))
sf8 = nv.Buf8(n, h, h, 16, halo=1)
sf8.t.copy_(torch.randint(120, 134, sf8.t.shape, device="cuda", dtype=torch.uint8))
w8 = torch.randint(0, 0x70, (c, 9 * c), device="cuda", dtype=torch.uint8)
wsf = torch.randint(120, 134, (c * 9 * c // 32,), device="cuda", dtype=torch.uint8)
out_mx = nv.Buf(n, h, h, c)
p = C.Plan()
mbw, mbh = nv.tile_shape(h, h)
no.add_conv_mx(p, "mxfp8", x8.tmap_padded(mbw, mbh), sf8.tmap_padded(mbw, mbh, cbox=16, swizzle=0),
               no.taps_conv(3, c), c, w8, wsf, out_mx.view(), h, h, n, c, stats=0 if a.no_stats else stats.data_ptr())
plans["mxfp8"] = (p, 2.0 * n * h * h * 9 * c * c)
p = C.Plan()   # tap-shared CTA-pair variant
no.add_conv_mx(p, "mxfp8_ts", x8.tmap_padded(8, 16), sf8.tmap_padded(8, 16, cbox=16, swizzle=0),
               no.taps_conv(3, c), c, w8, wsf, out_mx.view(), h, h, n, c, stats=0 if a.no_stats else stats.data_ptr(),
               tile=(8, 16), halo_maps=lambda w_, h_: (x8.tmap_padded(w_, h_), sf8.tmap_padded(w_, h_, cbox=16, swizzle=0)))
plans["mxfp8_ts"] = (p, 2.0 * n * h * h * 9 * c * c)
P = h + 2
rows = n * P * P
p = C.Plan()
taps = [(0, 0, (1 - r) * P + (1 - s), 0, (r * 3 + s) * c) for r in range(3) for s in range(3)]
no.add_conv_flat(p, "dgrad", dy.tmap_flat(), taps, c, w2, dxp.ptr(padded=True), c, rows, c, b_mn=True,
                 halo_map=(lambda rb: dy.tmap_flat(rows_box=rb)) if ts else None)
plans["dgrad"] = (p, 2.0 * n * h * h * 9 * c * c)
kbw, kbh = nv.tile_shape(h, h, 64)
p = C.Plan()
if True:
  no.add_wgrad(p, "wgrad", dy.tmap_interior(kbw, kbh), [x.tmap_padded(kbw, kbh)], no.taps_conv(3, c),
               h // kbw, h // kbh, kbw, kbh, n, c, c, dw.data_ptr(), 9 * c)
plans["wgrad"] = (p, 2.0 * n * h * h * 9 * c * c)

s = torch.cuda.current_stream().cuda_stream
res = {}
for name, (plan, flops) in plans.items():
  if a.only and a.only != name:
    continue
  for _ in range(3):
    plan.run(0, -1, s, False)
  torch.cuda.synchronize()
  ts = []
  for _ in range(a.iters):
    if not a.no_flush:
      flush.zero_(); flush2.sum()   # dirty then clean: the timed kernel does not pay the write-back
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    plan.run(0, -1, s, False)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
  ts.sort()
  ms = ts[len(ts) // 2]
  tf = flops / ms / 1e9
  res[name] = {"ms_median": ms, "ms_min": ts[0], "tflops": tf, "frac_of_measured_bf16_peak": tf / peak}
  print(f"{name:6s} images={n} median {ms * 1e3:8.1f} us  {tf:7.1f} TFLOP/s  "
        f"{100 * tf / peak:5.1f}% of measured cuBLAS bf16 peak ({peak:.0f})")
if a.mx_profile and "mxfp8_ts" in plans:
  prof = torch.zeros(16 + 5 * 40 + 4 * 14, dtype=torch.int64, device="cuda")
  C.set_mx_profile(prof.data_ptr())
  plans["mxfp8_ts"][0].run(0, -1, s, False)
  torch.cuda.synchronize()
  C.set_mx_profile(0)
  v = prof.tolist()
  names = ["producer total", "  wait a_empty", "  wait b_empty", "mma total", "  wait tmem_empty", "  wait a_full",
           "  wait b_full", "  wait sf_bar", "scale warp total", "  wait sfa_full", "  wait sfb_full", "epilogue warp total",
           "  wait tmem_full", "mma: 4 x tcgen05.mma", "mma: tcgen05.commit", "mma: elect..syncwarp block"]
  print("stall attribution, CTA 0 (cycles):")
  for nm, x in zip(names, v):
    print(f"  {nm:24s} {x:10d}")
  print("MMA warp time line (cycles since role start): stage start | b_full seen | sf_bar seen | MMAs issued | stage end")
  for i in range(40):
    row = v[16 + 5 * i:21 + 5 * i]
    if any(row):
      print("  stage %2d: " % i + " ".join(f"{x:8d}" for x in row))
  print("channel-block boundaries: block start | a_full seen | sf_bar seen | block end (a_empty committed)")
  for i in range(14):
    row = v[216 + 4 * i:220 + 4 * i]
    if any(row):
      print("  block %2d: " % i + " ".join(f"{x:8d}" for x in row))
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"shape": f"3x3 conv 256->256 @64x64 x{n} images", "l2_flush": not a.no_flush, "results": res},
          open("gpurun_out/kernel_bench.json", "w"), indent=1)
