"""Loss heads of the reference's CycleGAN class (main.py:172-195).

Every loss is ``sum_over_local_batch(per_sample_mean) / global_batch_size`` so
that the cross-replica gradient all-reduce is a plain SUM (main.py:172-174).
"""
from __future__ import annotations

import torch

from . import oracle as O

LAMBDA_CYCLE = 10.0                    # main.py:116
LAMBDA_IDENTITY = 0.5 * LAMBDA_CYCLE   # main.py:118


class LossHeads:

  def __init__(self, global_batch_size: int,
               lambda_cycle: float = LAMBDA_CYCLE,
               lambda_identity: float = LAMBDA_IDENTITY, lambda_adv: float = 1.0):
    self.global_batch_size = global_batch_size
    self.lambda_cycle = lambda_cycle
    self.lambda_identity = lambda_identity
    self.lambda_adv = lambda_adv       # 1 in the reference (main.py:176-179); a knob for per-term gradient tests

  def reduce_mean(self, per_sample: torch.Tensor) -> torch.Tensor:
    return per_sample.sum() / self.global_batch_size

  def generator_loss(self, discriminate_fake: torch.Tensor) -> torch.Tensor:
    return self.lambda_adv * self.reduce_mean(O.mse(torch.ones_like(discriminate_fake), discriminate_fake))

  def cycle_loss(self, real: torch.Tensor, cycled: torch.Tensor) -> torch.Tensor:
    return self.lambda_cycle * self.reduce_mean(O.mae(real, cycled))

  def identity_loss(self, real: torch.Tensor, same: torch.Tensor) -> torch.Tensor:
    return self.lambda_identity * self.reduce_mean(O.mae(real, same))

  def discriminator_loss(self, d_real: torch.Tensor, d_fake: torch.Tensor) -> torch.Tensor:
    real_loss = O.mse(torch.ones_like(d_real), d_real)
    fake_loss = O.mse(torch.zeros_like(d_fake), d_fake)
    return self.reduce_mean(0.5 * (real_loss + fake_loss))
