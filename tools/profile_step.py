"""Per-op device-time breakdown of one native train step (CUDA events around every plan op).

    python tools/profile_step.py [--batch 4] [--size 256] [--out gpurun_out/step_profile.txt]
"""
import argparse
import collections
import os
import re
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tf2_cyclegan_b200.engine.cyclegan import build_specs
from tf2_cyclegan_b200.engine.native_backend import NativeBackend
from tf2_cyclegan_b200.parallel.strategy import Strategy

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--nres", type=int, default=9)
ap.add_argument("--out", default="gpurun_out/step_profile.txt")
a = ap.parse_args()

st = Strategy(device="cuda", init_process_group=False)
be = NativeBackend(build_specs(types.SimpleNamespace(num_residual_blocks=a.nres)), a.batch, st,
                   image_size=a.size, local_batch=a.batch, use_graph=False)
x = torch.rand(a.batch, a.size, a.size, 3, device="cuda") * 2 - 1
y = torch.rand(a.batch, a.size, a.size, 3, device="cuda") * 2 - 1
for _ in range(3):
  be.train_step(x, y)
torch.cuda.synchronize()
sp = be.plan_for(a.batch)
s = torch.cuda.current_stream().cuda_stream
runs = [sp.plan.run_timed(0, sp.marks["end"], s) for _ in range(3)]
ms = [min(r[i] for r in runs) for i in range(len(runs[0]))]
names = sp.names
total = sum(ms)


def klass(n):
  n = re.sub(r"\[\d+:\d+\]", "", n)
  n = re.sub(r"\.p\d\d$", "", n)
  parts = n.split(".")
  if parts[0] in "GFXY" and len(parts) >= 3:
    layer = re.sub(r"\d+", "#", parts[1])
    return f"{'gen' if parts[0] in 'GF' else 'dis'}.{layer}.{parts[2]}"
  return parts[0] + "." + (parts[1] if len(parts) > 1 else "")


agg = collections.defaultdict(lambda: [0.0, 0])
for n, t in zip(names, ms):
  k = klass(n)
  agg[k][0] += t
  agg[k][1] += 1
lines = [f"native step profile: batch {a.batch}, {a.size}x{a.size}, {a.nres} res blocks; "
         f"sum of per-op event times {total:.3f} ms over {len(ms)} ops"]
lines.append(f"{'class':44s} {'ms':>9s} {'%':>6s} {'ops':>5s}")
for k, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
  lines.append(f"{k:44s} {t:9.3f} {100 * t / total:6.2f} {c:5d}")
lines.append("")
lines.append("top 40 single ops:")
for t, n in sorted(zip(ms, names), reverse=True)[:40]:
  lines.append(f"  {t:8.4f} ms  {n}")
os.makedirs(os.path.dirname(a.out), exist_ok=True)
open(a.out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:45]))
