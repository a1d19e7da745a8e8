#!/usr/bin/env python
"""Dependency critical path of the native train-step plan vs its serialised op time (CPU only).

Builds the step plan against the validating fake extension (tests/fake_native.py), takes per-op
device times from a `tools/profile_step.py` report (class totals, split over the ops of a class in
proportion to their image counts) and propagates earliest-finish times through the plan's
dependency structure: in-order lanes, fork / join, event record / wait.

    python tools/critical_path.py profiles/r1_step_profile_v8.txt [--batch 4] [--size 256]

critical path ~ step time  -> the step is dependency (latency) bound: more lanes / fusion help;
critical path << step time -> the step is throughput bound: only less work per op helps.
"""
import argparse
import collections
import os
import re
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pytest
import torch

from tests import fake_native

ap = argparse.ArgumentParser()
ap.add_argument("profile")
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--nres", type=int, default=9)
a = ap.parse_args()

mp = pytest.MonkeyPatch()
fake_native.install(mp)
mp.setattr(torch.cuda, "synchronize", lambda *x, **k: None)
from tf2_cyclegan_b200.engine.cyclegan import build_specs  # noqa: E402
from tf2_cyclegan_b200.engine.native_backend import NativeBackend  # noqa: E402
from tf2_cyclegan_b200.parallel.strategy import Strategy  # noqa: E402

plan_cls = next(v for v in vars(fake_native).values() if isinstance(v, type) and hasattr(v, "_add"))
lanes = []
orig = plan_cls._add


def _add(self, kind, name, *args):
  lanes.append(getattr(self, "lane", 0))
  return orig(self, kind, name, *args)


plan_cls._add = _add
be = NativeBackend(build_specs(types.SimpleNamespace(num_residual_blocks=a.nres)), a.batch,
                   Strategy(device="cpu"), seed=1, image_size=a.size, local_batch=a.batch)
sp = be.plan_for(a.batch)
ops = sp.plan.ops[:sp.marks["end"]]
lanes = lanes[:len(ops)]


def klass(n):
  n = re.sub(r"\[\d+:\d+\]", "", n)
  n = re.sub(r"\.p\d\d$", "", n)
  parts = n.split(".")
  if parts[0] in "GFXY" and len(parts) >= 3:
    layer = re.sub(r"\d+", "#", parts[1])
    return f"{'gen' if parts[0] in 'GF' else 'dis'}.{layer}.{parts[2]}"
  return parts[0] + "." + (parts[1] if len(parts) > 1 else "")


def nimg(n):
  m = re.search(r"\[(\d+):(\d+)\]", n)
  return int(m.group(2)) - int(m.group(1)) if m else 1


cls_ms = {}
for line in open(a.profile):
  m = re.match(r"(\S+)\s+([\d.]+)\s+[\d.]+\s+(\d+)\s*$", line)
  if m:
    cls_ms[m.group(1)] = float(m.group(2))
weight = collections.defaultdict(float)
for kind, name, _ in ops:
  weight[klass(name)] += nimg(name)
dur = [cls_ms.get(klass(name), 0.0) * nimg(name) / weight[klass(name)] if kind not in
       ("fork", "join", "record", "wait") else 0.0 for kind, name, _ in ops]

finish = [0.0] * len(ops)
lane_last = [None, None, None, None]     # index of the last op issued on each lane
forked, fork_at = False, None
for i, (kind, name, args) in enumerate(ops):
  lane = lanes[i] if forked else 0
  deps = []
  if kind == "fork":
    forked, fork_at = True, lane_last[0]
    for k in range(1, 4):
      lane_last[k] = fork_at
    finish[i] = finish[fork_at] if fork_at is not None else 0.0
    lane_last[0] = i
    continue
  if kind == "join":
    deps = [x for x in lane_last if x is not None]
    finish[i] = max([finish[x] for x in deps] or [0.0])
    forked = False
    lane_last = [i, None, None, None]
    continue
  if lane_last[lane] is not None:
    deps.append(lane_last[lane])
  if kind == "wait":
    deps.append(args[1])                   # the record op
  finish[i] = max([finish[x] for x in deps] or [0.0]) + dur[i]
  lane_last[lane] = i

total = sum(dur)
cp = max(finish)
busy = collections.defaultdict(float)
for i, d in enumerate(dur):
  busy[lanes[i]] += d
print(f"ops {len(ops)}; serialised op time {total:.2f} ms; dependency critical path {cp:.2f} ms")
print("per-lane op time: " + ", ".join(f"lane {k}: {v:.2f} ms" for k, v in sorted(busy.items())))
# the chain itself
i = max(range(len(ops)), key=lambda k: finish[k])
print(f"last op on the critical path: {ops[i][1]}")
