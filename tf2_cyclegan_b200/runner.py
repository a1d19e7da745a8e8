"""Run orchestration: the reference's ``main`` / ``train`` / ``test``
(main.py:332-402) on top of the B200 engine.

Same behaviour: optional output-dir wipe, seeds 1234, ``global_batch_size =
replicas × batch_size``, auto-resume, per-epoch train → test → ``elapse``
scalar → MAE print-out (including the reference's swapped labels,
main.py:394-398) → checkpoint + cycle plots every 10 epochs and at the end.
Differences: one process per GPU (only the chief writes files); loss scalars
stay on the device until the epoch ends (one D2H per epoch instead of one per
step); the epoch counter is additionally saved in the checkpoint but — like
the reference — a resumed run still counts epochs from 0 unless
``--resume_epoch`` is given.
"""
from __future__ import annotations

import os
import random
from shutil import rmtree
from time import time

import numpy as np
import torch
from tqdm import tqdm

from .data.pipeline import get_datasets
from .engine.cyclegan import CycleGAN
from .parallel.strategy import Strategy
from .utils import plot
from .utils.dicts import append_dict
from .utils.summary import Summary
from .utils.trace import nvtx_range


def _epoch_means(results: dict) -> dict:
  """mean over steps of the all-replica SUM, one host transfer for all keys."""
  if not results:
    return {}
  keys = list(results.keys())
  stacked = torch.stack([torch.stack([v.detach().float().reshape(()) for v in results[k]]).mean()
                         for k in keys])
  vals = stacked.cpu().tolist()
  return dict(zip(keys, vals))


def _check_finite(means: dict, where: str) -> None:
  """Failure detection the reference lacks: stop on the first non-finite epoch mean instead of
  silently training on NaNs (checked once per epoch, on values already on the host)."""
  bad = [k for k, v in means.items() if v != v or v in (float("inf"), float("-inf"))]
  if bad:
    raise FloatingPointError(f"non-finite {where} metrics: {bad}")


def train(args, train_ds, gan, summary, epoch: int):
  results = {}
  dev = gan.strategy.device
  timed = dev.type == "cuda"
  if timed:   # device-side step timing (CUDA events), logged as `step_ms`
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
  steps = 0
  for x, y in tqdm(train_ds, desc="Train", total=args.train_steps,
                   disable=args.verbose == 0 or not gan.strategy.is_chief):
    with nvtx_range("train_step"):
      result = gan.distributed_train_step(x, y)
    append_dict(results, result)
    steps += 1
  if timed:
    e1.record()
  means = _epoch_means(results)
  _check_finite(means, "train")
  for key, value in means.items():
    summary.scalar(key, value, step=epoch, training=True)
  if timed and steps:
    e1.synchronize()
    summary.scalar("step_ms", e0.elapsed_time(e1) / steps, step=epoch, training=True)


def test(args, test_ds, gan, summary, epoch: int):
  results = {}
  for x, y in tqdm(test_ds, desc="Test", total=args.test_steps,
                   disable=args.verbose == 0 or not gan.strategy.is_chief):
    with nvtx_range("test_step"):
      result = gan.distributed_test_step(x, y)
    append_dict(results, result)
  results = _epoch_means(results)
  _check_finite(results, "test")
  for key, value in results.items():
    summary.scalar(key, value, step=epoch, training=False)
  return results


def main(args, strategy: Strategy | None = None):
  own_strategy = strategy is None
  if strategy is None:
    strategy = Strategy(device=getattr(args, "device", None))
  if strategy.is_chief:
    if args.clear_output_dir and os.path.exists(args.output_dir):
      rmtree(args.output_dir)
    os.makedirs(args.output_dir, exist_ok=True)
  strategy.barrier()

  seed = int(getattr(args, "seed", 1234))
  random.seed(seed)
  np.random.seed(seed)
  torch.manual_seed(seed)

  num_devices = strategy.num_replicas_in_sync
  args.global_batch_size = num_devices * args.batch_size
  if strategy.is_chief:
    print(f"Number of devices: {num_devices}")

  summary = Summary(args.output_dir, enabled=strategy.is_chief)
  train_ds, test_ds, plot_ds = get_datasets(args, strategy=strategy)
  gan = CycleGAN(args, strategy=strategy)

  # load last checkpoint if exists
  gan.load_checkpoint()
  start_epoch = 0
  if getattr(args, "resume_epoch", False):
    start_epoch = int(gan.extra_state.get("epoch", -1)) + 1

  results = {}
  for epoch in range(start_epoch, args.epochs):
    if strategy.is_chief:
      print(f"Epoch {epoch + 1:03d}/{args.epochs:03d}")

    start = time()
    train(args, train_ds, gan, summary, epoch)
    results = test(args, test_ds, gan, summary, epoch)
    end = time()
    summary.scalar("elapse", end - start, step=epoch)

    if strategy.is_chief:
      # NOTE: the reference prints the 2nd and 3rd values under swapped labels
      # (main.py:394-398); kept verbatim for log parity.
      print(f'MAE(X, F(G(X))): {results["error/MAE(X, F(G(X)))"]:.04f}\t\t'
            f'MAE(X, F(X)): {results["error/MAE(Y, G(F(Y)))"]:.04f}\n'
            f'MAE(Y, G(F(Y))): {results["error/MAE(X, F(X))"]:.04f}\t\t'
            f'MAE(Y, G(Y)): {results["error/MAE(Y, G(Y))"]:.04f}\n'
            f'Elapse: {end - start:.02f}s\n')

    if epoch % 10 == 0 or epoch == args.epochs - 1:
      gan.save_checkpoint(extra={"epoch": epoch, "seed": seed})
      if strategy.is_chief:
        plot.plot_cycle(plot_ds, gan, summary, epoch)
      strategy.barrier()

  summary.close()
  if own_strategy:
    strategy.shutdown()
  return results
