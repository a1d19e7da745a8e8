"""``CycleGAN``: the training-engine facade with the reference's public surface
(main.py:106-329): ``distributed_train_step`` / ``distributed_test_step`` /
``cycle_step`` / ``save_checkpoint`` / ``load_checkpoint(expect_partial)``.

Two interchangeable back-ends implement the step:
  * ``oracle`` — PyTorch autograd (CPU or CUDA eager); ground truth.
  * ``native`` — hand-scheduled sm_100a kernels, flat parameter buffers, fused
    peer-memory gradient reduce + Adam (engine/native_backend.py).
"""
from __future__ import annotations

import os
from typing import Dict

import torch

from ..models.spec import discriminator_spec, generator_spec
from ..parallel.strategy import Strategy
from ..utils import checkpoint as ckpt


def build_specs(args) -> Dict[str, object]:
  nres = int(getattr(args, "num_residual_blocks", 9))
  return {"G": generator_spec(num_residual_blocks=nres, name="gen_G"),
          "F": generator_spec(num_residual_blocks=nres, name="gen_F"),
          "X": discriminator_spec(name="dis_X"),
          "Y": discriminator_spec(name="dis_Y")}


class CycleGAN:
  """CycleGAN class that supports per replica loss calculation."""

  def __init__(self, args, strategy: Strategy):
    self.strategy = strategy
    self.global_batch_size = args.global_batch_size
    self.checkpoint_format = getattr(args, "checkpoint_format", "native")
    self.lambda_cycle = 10.0
    self.lambda_identity = 0.5 * self.lambda_cycle

    self.checkpoint_dir = os.path.join(args.output_dir, "checkpoints")
    if strategy.is_chief:
      os.makedirs(self.checkpoint_dir, exist_ok=True)
    self.checkpoint_prefix = os.path.join(self.checkpoint_dir, "checkpoint")

    self.specs = build_specs(args)
    backend = getattr(args, "backend", "auto")
    dtype = getattr(args, "dtype", "fp32")
    seed = int(getattr(args, "seed", 1234))
    size = int(getattr(args, "image_size", 256))
    if backend == "auto":
      native_ok = strategy.device.type == "cuda" and size >= 32 and size % 8 == 0
      backend = "native" if native_ok else "oracle"
    emulate_mx = backend == "oracle" and dtype == "mxfp8"   # oracle: fp32 math on MXFP8-quantised operands
    if backend == "oracle" and dtype in ("fp8", "mxfp8"):
      dtype = "fp32" if emulate_mx else "bf16"
    if backend == "native":
      from .native_backend import NativeBackend
      self.backend = NativeBackend(self.specs, self.global_batch_size, strategy, seed=seed,
                                   image_size=int(getattr(args, "image_size", 256)),
                                   local_batch=int(args.batch_size), dtype=dtype,
                                   fused_allreduce=getattr(args, "fused_allreduce", "auto"))
    elif backend == "oracle":
      from .oracle_backend import OracleBackend
      cd = {"fp32": torch.float32, "bf16": torch.bfloat16}[dtype]
      self.backend = OracleBackend(self.specs, self.global_batch_size, strategy, seed=seed,
                                   compute_dtype=cd, emulate_mxfp8=emulate_mx)
    else:
      raise ValueError(f"unknown backend {backend!r}")
    self.extra_state: Dict[str, object] = {}

  # ---- checkpoints (main.py:157-170) ------------------------------------
  def save_checkpoint(self, extra: Dict[str, object] | None = None):
    """save checkpoint to checkpoint_dir, overwrite if exists (chief only)."""
    sd = self.backend.state_dict()  # collective for sharded optimizer state
    more = getattr(self.backend, "extra_state", lambda: {})()
    if more:
      extra = dict(extra or {}, backend=more)
    if self.strategy.is_chief:
      ckpt.save(self.checkpoint_prefix, sd, extra=extra,
                fmt=self.checkpoint_format)
      print(f"\nsaved checkpoint to {self.checkpoint_prefix}\n")
    self.strategy.barrier()

  def load_checkpoint(self, expect_partial: bool = False) -> bool:
    """load checkpoint from checkpoint_dir if exists."""
    if ckpt.exists(self.checkpoint_prefix):
      sd, extra = ckpt.load(self.checkpoint_prefix)
      self.backend.load_state_dict(sd, expect_partial=expect_partial)
      if hasattr(self.backend, "load_extra_state"):
        self.backend.load_extra_state((extra or {}).get("backend", {}))
      self.extra_state = extra
      if self.strategy.is_chief:
        print(f"\nloaded checkpoint from {self.checkpoint_prefix}\n")
      return True
    return False

  # ---- steps -------------------------------------------------------------
  def cycle_step(self, x, y, training: bool = False):
    """x -> fake y -> cycle x ; y -> fake x -> cycle y (main.py:197-205)."""
    return self.backend.cycle_step(x, y)

  def train_step(self, x, y):
    return self.backend.train_step(x, y)

  def test_step(self, x, y):
    return self.backend.test_step(x, y)

  def reduce_dict(self, d: dict):
    """SUM-reduce result scalars over replicas (main.py:264-267). The native fused path has already
    summed them on the device (loss-scalar side channel of the last fused reduce + Adam launch): no
    collective call and no host synchronisation per step."""
    reduce = getattr(self.backend, "reduce_results", None)
    return reduce(d) if reduce is not None else self.strategy.reduce_dict(d)

  def distributed_train_step(self, x, y):
    return self.reduce_dict(self.train_step(x, y))

  def distributed_test_step(self, x, y):
    return self.reduce_dict(self.test_step(x, y))
