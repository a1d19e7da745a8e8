#!/usr/bin/env python
"""Headline benchmark: CycleGAN train-step throughput (images/sec at 256x256).

    python bench.py --gpus N --steps K --warmup W [--impl native|standin|oracle|reference]

* metric / config are BASELINE.json's: 256x256, 9-block ResNet generators + PatchGAN
  discriminators, bf16 compute, batch 4 per GPU, synthetic data, random-init weights.
  One "image" = one (x, y) pair slot of --batch_size (main.py:71-72,372).
* `value` = whole-job images/sec, device-timed with CUDA events between a barrier +
  synchronize on both sides, max over ranks. The timed region is the full train step:
  6 generator + 4 discriminator forwards, all backward passes, 4 gradient
  reductions and 4 Adam updates.
* `e2e` = the same metric through the public API (CycleGAN.distributed_train_step)
  including the per-step H2D copy of the inputs from pinned memory and the D2H read
  of the loss.
* N > 1: launched by torchrun (one rank per GPU); rank 0 prints the JSON line.
* `--impl reference`: the unmodified TF-2.4 reference cannot be installed offline on
  this image -> prints {"impl": "reference", "unavailable": ...}.
* `--impl standin`: baseline/standin.py — the strongest stock-PyTorch implementation of the same
  step (channels_last bf16 + cuDNN + ATen instance norm + flat fp32 gradient buffers with one packed
  NCCL all-reduce per network + multi-tensor Keras-Adam, the whole step in one CUDA graph). It is a
  STAND-IN, not the reference; same metric, config, timing rules and JSON line.
* `--impl oracle`: the fp32-semantics autograd oracle of the test-suite run under autocast (eager,
  unfused); context only, not a tuned baseline.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def reference_arm():
  why = None
  ref_dir = os.path.join(ROOT, "baseline", "_ref")
  try:
    sys.path.insert(0, ref_dir)
    import tensorflow  # noqa: F401  (reference imports tensorflow 2.4.1 + tfa + tfds)
  except Exception as e:  # pragma: no cover
    why = (f"reference needs tensorflow==2.4.1/tensorflow_addons/tensorflow_datasets "
           f"(python<=3.8, CUDA 11.0); not in the offline wheelhouse: {type(e).__name__}: {e}")
  finally:
    if ref_dir in sys.path:
      sys.path.remove(ref_dir)
  if why is None:
    why = "tensorflow import succeeded unexpectedly but the reference harness is not wired"
  print(json.dumps({"impl": "reference", "unavailable": why.replace("\n", " ")[:400]}))
  return 0


def standin_arm(a, strategy, ClockSampler):
  """baseline/standin.py (stock PyTorch library kernels; NOT the reference), same protocol as main()."""
  import torch
  sys.path.insert(0, os.path.join(ROOT, "baseline"))
  import standin
  torch.backends.cudnn.benchmark = True
  dev, world = strategy.device, strategy.world_size
  gbs, S = a.batch_size * world, a.image_size
  m = standin.StandinCycleGAN(dev, a.batch_size, gbs, image_size=S, n_res=a.num_residual_blocks, world=world)
  graphed = m.capture()
  g = torch.Generator().manual_seed(1234 + strategy.rank)
  host = [(torch.rand(a.batch_size, S, S, 3, generator=g).mul_(2).sub_(1).pin_memory(),
           torch.rand(a.batch_size, S, S, 3, generator=g).mul_(2).sub_(1).pin_memory()) for _ in range(4)]
  xd, yd = host[0][0].to(dev), host[0][1].to(dev)

  def timed(fn, steps):
    strategy.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
      fn(i)
    e1.record(); torch.cuda.synchronize(); strategy.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    strategy.all_reduce_max(ms)
    return ms.item()

  def e2e_step(i):
    x, y = host[i % 4]
    return float(m.train_step(x.to(dev, non_blocking=True), y.to(dev, non_blocking=True))[3])   # D2H loss_G/total

  for _ in range(max(a.warmup, 3)):
    m.train_step(xd, yd)
  sampler = ClockSampler(strategy.local_rank); sampler.start()
  ms = timed(lambda i: m.train_step(xd, yd), a.steps)
  clocks = sampler.stop()
  for i in range(3):
    e2e_step(i)
  ms_e = timed(e2e_step, a.steps)
  if strategy.is_chief:
    print(json.dumps({
        "metric": "images/sec at 256x256 (train step, whole job, device-timed, max over ranks)",
        "value": gbs * a.steps / (ms / 1e3), "unit": "images/sec", "n_gpus": world, "steps": a.steps,
        "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "impl": "standin",
        "note": "stand-in baseline (stock PyTorch + cuDNN + NCCL), NOT the reference: baseline/standin.py",
        "config": {"model": f"CycleGAN {S}x{S}, {a.num_residual_blocks}-block ResNet G/F + 70x70 PatchGAN X/Y",
                   "global_batch": gbs, "batch_per_gpu": a.batch_size, "seq_len": S * S, "parallelism": f"dp{world}",
                   "cuda_graph": graphed, "collective": "nccl" if world > 1 else "none",
                   "l2": "no explicit flush: per-step activation working set (>2 GB) exceeds the 126 MB L2"},
        "clocks": clocks,
        "e2e": {"value": gbs * a.steps / (ms_e / 1e3), "unit": "images/sec", "ms_per_step": ms_e / a.steps,
                "h2d_bytes_per_step": 2 * a.batch_size * S * S * 3 * 4, "d2h_bytes_per_step": 4,
                "api": "StandinCycleGAN.train_step(x_pinned.to(cuda), y_pinned.to(cuda)) + float(loss)"}}))
  strategy.barrier()
  strategy.shutdown()
  return 0


class ClockSampler:
  """Samples SM clock / throttle reasons through NVML during the timed region."""

  def __init__(self, index: int):
    self.samples, self.reasons, self.max_mhz = [], set(), None
    self._stop = threading.Event()
    self._t = None
    try:
      import pynvml
      pynvml.nvmlInit()
      self.nv = pynvml
      self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
      self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
    except Exception:
      self.nv = None

  def _loop(self):
    nv = self.nv
    names = {
        nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
        nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
        nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
        nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
    }
    while not self._stop.is_set():
      try:
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for bit, name in names.items():
          if mask & bit:
            self.reasons.add(name)
      except Exception:
        pass
      self._stop.wait(0.05)

  def start(self):
    if self.nv is not None:
      self._t = threading.Thread(target=self._loop, daemon=True)
      self._t.start()

  def stop(self):
    self._stop.set()
    if self._t is not None:
      self._t.join()
    s = sorted(self.samples)
    return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons), "samples": len(s)}


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=20)
  ap.add_argument("--warmup", type=int, default=5)
  ap.add_argument("--impl", default="native", choices=["native", "standin", "oracle", "reference"])
  ap.add_argument("--batch_size", type=int, default=4, help="per GPU (BASELINE config 2: 4)")
  ap.add_argument("--image_size", type=int, default=256)
  ap.add_argument("--num_residual_blocks", type=int, default=9)
  ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32", "fp8", "mxfp8"])
  ap.add_argument("--fused_allreduce", default="on", choices=["auto", "on", "off"],
                  help="gradient path on > 1 GPU: 'on' (default) = the fused peer-memory reduce + Adam kernel or "
                       "fail loudly; 'off' = packed NCCL all-reduce + Adam kernel (baseline path)")
  ap.add_argument("--ar_mode", default=None, choices=["push", "pull"],
                  help="form of the fused gradient reduction (default: parallel.fused_allreduce.DEFAULT_MODE)")
  ap.add_argument("--no_e2e", action="store_true")
  a = ap.parse_args()
  if a.impl == "reference":
    return reference_arm()

  import torch
  from tf2_cyclegan_b200.engine.cyclegan import CycleGAN
  from tf2_cyclegan_b200.parallel.strategy import Strategy

  if not torch.cuda.is_available():
    print(json.dumps({"error": "bench.py needs a CUDA device"}))
    return 1
  world = int(os.environ.get("WORLD_SIZE", "1"))
  if world != a.gpus:
    if world == 1 and a.gpus > 1:
      # convenience: re-launch ourselves under torchrun
      import subprocess
      cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
             f"--nproc-per-node={a.gpus}", "--master-addr", "127.0.0.1", "--master-port", "29533",
             os.path.abspath(__file__)] + sys.argv[1:]
      return subprocess.call(cmd)
    raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}")

  if a.ar_mode:
    from tf2_cyclegan_b200.config import ENGINE
    from tf2_cyclegan_b200.parallel import fused_allreduce
    ENGINE.ar_mode = fused_allreduce.DEFAULT_MODE = a.ar_mode
  strategy = Strategy(device="cuda")
  dev = strategy.device
  gbs = a.batch_size * world
  if a.impl == "standin":
    return standin_arm(a, strategy, ClockSampler)
  args = types.SimpleNamespace(output_dir=os.path.join("/tmp", f"bench_{os.getpid()}"),
                               global_batch_size=gbs, batch_size=a.batch_size,
                               image_size=a.image_size, num_residual_blocks=a.num_residual_blocks,
                               backend=a.impl, dtype=a.dtype, seed=1234,
                               fused_allreduce=a.fused_allreduce)
  gan = CycleGAN(args, strategy)
  be = gan.backend
  S = a.image_size
  g = torch.Generator().manual_seed(1234 + strategy.rank)
  n_host = 4  # rotate through a few distinct pinned host batches
  host = [(torch.rand(a.batch_size, S, S, 3, generator=g).mul_(2).sub_(1).pin_memory(),
           torch.rand(a.batch_size, S, S, 3, generator=g).mul_(2).sub_(1).pin_memory())
          for _ in range(n_host)]
  xd, yd = host[0][0].to(dev), host[0][1].to(dev)

  def device_step():
    return be.train_step(xd, yd)

  def e2e_step(i):
    x, y = host[i % n_host]
    res = gan.distributed_train_step(x.to(dev, non_blocking=True), y.to(dev, non_blocking=True))
    return float(res["loss_G/total"])  # D2H read of the step's result

  def timed(fn, steps, with_index=False):
    strategy.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
      fn(i) if with_index else fn()
    e1.record()
    torch.cuda.synchronize()
    strategy.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    strategy.all_reduce_max(ms)
    return ms.item()

  launches0 = getattr(be, "kernel_launches", lambda: 0)()
  for _ in range(max(a.warmup, 3)):
    device_step()
  torch.cuda.synchronize()
  launches_per_step = (getattr(be, "kernel_launches", lambda: 0)() - launches0) // max(a.warmup, 3)

  sampler = ClockSampler(strategy.local_rank)
  sampler.start()
  ms = timed(device_step, a.steps)
  clocks = sampler.stop()
  value = gbs * a.steps / (ms / 1e3)

  e2e = None
  if not a.no_e2e:
    api = None
    if a.impl == "native":
      # the repo's own input pipeline (what runner.train iterates): host dataset -> native batch
      # gather into rotating pinned staging slots -> async H2D on a side stream, two batches
      # ahead -> train step -> D2H read of the loss, every step
      try:
        from tf2_cyclegan_b200.data.pipeline import PairedBatches
        warm, ahead = 3, 2
        n_steps = warm + a.steps + ahead      # `ahead` extra batches: the prefetcher keeps uploading
        nb = a.batch_size                     # every rank feeds its own local batches
        base = torch.rand(8 * nb, S, S, 3, generator=g).mul_(2).sub_(1)
        data_a = base.repeat((n_steps * nb + len(base) - 1) // len(base), 1, 1, 1)[:n_steps * nb].contiguous()
        data_b = data_a.flip(0).contiguous()
        it = iter(PairedBatches(data_a, data_b, nb, strategy, shuffle_buffer=0, distribute=False,
                                prefetch=ahead))

        def pipe_step():
          x, y = next(it)
          return float(gan.distributed_train_step(x, y)["loss_G/total"])

        for _ in range(warm):
          pipe_step()
        ms_e = timed(pipe_step, a.steps)
        it.close()                            # stops and joins the prefetch thread
        api = ("for x, y in PairedBatches(host tensors, prefetch=2): float(CycleGAN.distributed_train_step(x, y)"
               "['loss_G/total'])  # pinned staging + side-stream H2D of every batch, D2H loss every step")
      except Exception as exc:  # fall back to the plain copy-then-step loop
        print(f"[bench] pipeline e2e failed ({type(exc).__name__}: {exc}); using the serial loop", file=sys.stderr)
        api = None
    if api is None:
      for i in range(3):
        e2e_step(i)
      ms_e = timed(e2e_step, a.steps, with_index=True)
      api = "CycleGAN.distributed_train_step(x_pinned.to(cuda), y_pinned.to(cuda)) + float(loss)"
    e2e = {"value": gbs * a.steps / (ms_e / 1e3), "unit": "images/sec",
           "ms_per_step": ms_e / a.steps,
           "h2d_bytes_per_step": 2 * a.batch_size * S * S * 3 * 4,
           "d2h_bytes_per_step": 4,
           "api": api}

  if strategy.is_chief:
    fwd = 6 * gan.specs["G"].forward_flops(S, S) + 6 * gan.specs["X"].forward_flops(S, S)
    out = {
        "metric": "images/sec at 256x256 (train step, whole job, device-timed, max over ranks)",
        "value": value, "unit": "images/sec", "n_gpus": world, "steps": a.steps,
        "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": a.dtype if a.impl == "native" else
        ("bf16 autocast" if a.dtype == "bf16" else a.dtype), "data": "synthetic",
        "impl": a.impl,
        "config": {"model": f"CycleGAN {S}x{S}, {a.num_residual_blocks}-block ResNet G/F + 70x70 PatchGAN X/Y",
                   "global_batch": gbs, "batch_per_gpu": a.batch_size, "seq_len": S * S,
                   "parallelism": f"dp{world}",
                   "l2": "no explicit flush: per-step activation working set (>2 GB) exceeds the 126 MB L2",
                   "pairs_per_sec": value, "total_images_per_sec": 2 * value,
                   "approx_train_tflop_per_pair": 3.0 * fwd / 1e12},
        "clocks": clocks, "gpu_launches": int(launches_per_step) * a.steps,
        "gpu_launches_per_step": int(launches_per_step),
    }
    if a.impl == "native":
      names = be.plan_for(a.batch_size).names
      out["collective"] = be.collective()              # fused-<push|pull>-<nvls|p2p> | nccl | none
      out["allreduce_adam_launches_per_step"] = sum(1 for n in names if "fused_allreduce_adam" in n)
      out["nccl_calls_per_step"] = 0 if (world == 1 or be.fused) else 5   # 4 gradient + 1 scalar all-reduce
      out["config"]["collective"] = out["collective"]
    if e2e is not None:
      out["e2e"] = e2e
    print(json.dumps(out))
  strategy.barrier()
  strategy.shutdown()
  return 0


if __name__ == "__main__":
  sys.exit(main())
