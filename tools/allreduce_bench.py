"""Gradient all-reduce sweep (BASELINE.json config 5): fused reduce+Adam+broadcast kernel vs
NCCL all_reduce (+ the separate Adam / cast kernels it needs), 1 KB - 1 GB, device-timed, max
over ranks. Run under torchrun.   bus GB/s = S * 2(N-1)/N / t   (fraction of 900 GB/s/dir)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from tf2_cyclegan_b200.ops import native as nv
from tf2_cyclegan_b200.parallel.fused_allreduce import DEFAULT_MODE, SymmetricBuffers, default_blocks
from tf2_cyclegan_b200.parallel.strategy import Strategy

st = Strategy(device="cuda")
C = nv.load()
sb = SymmetricBuffers(st)
W = st.world_size
sizes = [1 << k for k in range(10, 31, 2)] + [45_533_708 // 32 * 32, 11_062_532 // 32 * 32]
if "--quick" in sys.argv:      # the sizes the train step uses (+ 1 MB, 4 MB)
  sizes = [1 << 20, 1 << 22, 11_062_532 // 32 * 32, 45_533_708 // 32 * 32]
rows = []


def timed(fn, iters):
  for _ in range(3):
    fn()
  st.barrier(); torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(iters):
    fn()
  e1.record(); torch.cuda.synchronize()
  t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
  st.all_reduce_max(t)
  return t.item()


for nbytes in sorted(sizes):
  n = max(8 * W, nbytes // 4 // (8 * W) * (8 * W))
  g, wb = sb.alloc(f"s{nbytes}", n)
  class S: pass
  s = S(); s.n = n
  s.w = torch.randn(n, device="cuda"); s.m = torch.zeros(n, device="cuda"); s.v = torch.zeros(n, device="cuda")
  s.step = torch.zeros(1, dtype=torch.int32, device="cuda")
  stream = torch.cuda.current_stream().cuda_stream
  iters = 50 if nbytes < (1 << 24) else 10
  t_mode = {}
  for mode in ("push", "pull"):
    for exposed in (False, True):
      plan = C.Plan()
      sb.emit_allreduce_adam(plan, f"s{nbytes}", s, 2e-4, 0.5, 0.9, 1e-7, mode=mode, exposed=exposed)
      t_mode[(mode, exposed)] = timed(lambda: plan.run(0, -1, stream, False), iters)
  t_fused = t_mode[(DEFAULT_MODE, True)]          # a stand-alone launch is an exposed one
  grid_sweep = {}
  if n * 4 in (11_062_528 // (32 * W) * (32 * W), 45_533_696 // (32 * W) * (32 * W)) or 11_000_000 < n * 4 < 11_100_000 or 45_500_000 < n * 4 < 45_600_000:
    for blocks in (2, 4, 8, 16, 32, 64):
      plan = C.Plan()
      sb.emit_allreduce_adam(plan, f"s{nbytes}", s, 2e-4, 0.5, 0.9, 1e-7, mode="push", blocks=blocks)
      grid_sweep[blocks] = timed(lambda: plan.run(0, -1, stream, False), iters)
  # NCCL baseline: packed fp32 all_reduce, then Adam + bf16 cast + grad memset kernels
  gb = torch.randn(n, device="cuda")
  p2 = C.Plan()
  wb2 = torch.zeros(n, device="cuda", dtype=torch.bfloat16)
  p2.add_adam("adam", s.w.data_ptr(), s.m.data_ptr(), s.v.data_ptr(), gb.data_ptr(), wb2.data_ptr(),
              s.step.data_ptr(), 2e-4, 0.5, 0.9, 1e-7, 1.0, n)
  p2.add_memset("zero", gb.data_ptr(), n * 4)
  t_nccl_only = timed(lambda: dist.all_reduce(gb), iters)
  t_nccl = timed(lambda: (dist.all_reduce(gb), p2.run(0, -1, stream, False)), iters)
  bus = lambda t: (n * 4) * 2 * (W - 1) / W / (t * 1e-3) / 1e9
  rows.append({"bytes": n * 4, "fused_ms": t_fused,
               "push_ms": t_mode[("push", True)], "pull_ms": t_mode[("pull", True)],
               "push_small_grid_ms": t_mode[("push", False)], "pull_small_grid_ms": t_mode[("pull", False)],
               "blocks": {"small": default_blocks(n * 4, False, W), "exposed": default_blocks(n * 4, True, W)},
               "push_grid_sweep_ms": grid_sweep,
               "nccl_allreduce_ms": t_nccl_only,
               "nccl_plus_adam_ms": t_nccl, "fused_bus_GBs": bus(t_fused),
               "nccl_bus_GBs": bus(t_nccl_only), "fused_frac_of_900": bus(t_fused) / 900.0})
  del g, wb
if st.is_chief:
  for r in rows:
    print(f"{r['bytes']:>11d} B  push {r['push_ms'] * 1e3:7.1f} us  pull {r['pull_ms'] * 1e3:7.1f} us  "
          f"(small grid {r['push_small_grid_ms'] * 1e3:6.1f} / {r['pull_small_grid_ms'] * 1e3:6.1f})  "
          f"nccl {r['nccl_allreduce_ms'] * 1e3:7.1f}  nccl+adam {r['nccl_plus_adam_ms'] * 1e3:7.1f} us  "
          f"fused bus {r['fused_bus_GBs']:6.1f} GB/s ({100 * r['fused_frac_of_900']:.1f}% of 900)")
    if r["push_grid_sweep_ms"]:
      print("            push by CTAs: " + "  ".join(f"{b}: {t * 1e3:.1f} us" for b, t in r["push_grid_sweep_ms"].items()))
  out = {"world": W, "mode": DEFAULT_MODE, "multicast": sb.multicast_ptrs(f"s{sorted(sizes)[0]}")[0] != 0, "rows": rows}
  print(json.dumps(out))
  os.makedirs("gpurun_out", exist_ok=True)
  json.dump(out, open(f"gpurun_out/allreduce_sweep_w{W}.json", "w"), indent=1)
st.barrier(); st.shutdown()
