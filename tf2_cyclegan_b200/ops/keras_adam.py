"""tf.keras.optimizers.Adam (TF 2.4) update rule, as used at main.py:134-145.

    lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t)
    m   <- b1 * m + (1 - b1) * g
    v   <- b2 * v + (1 - b2) * g^2
    w   <- w - lr_t * m / (sqrt(v) + eps)          # eps OUTSIDE the bias correction

Defaults are the reference's: lr 2e-4, beta_1 0.5, beta_2 0.9 (not 0.999),
eps 1e-7 (Keras default). This is the oracle for the fused reduce+Adam kernel.
"""
from __future__ import annotations

import math
from typing import Dict

import torch


class KerasAdam:

  def __init__(self, lr: float = 2e-4, beta_1: float = 0.5, beta_2: float = 0.9,
               epsilon: float = 1e-7):
    self.lr, self.beta_1, self.beta_2, self.epsilon = lr, beta_1, beta_2, epsilon
    self.iterations = 0
    self.m: Dict[str, torch.Tensor] = {}
    self.v: Dict[str, torch.Tensor] = {}

  def step_size(self, t: int) -> float:
    return self.lr * math.sqrt(1.0 - self.beta_2 ** t) / (1.0 - self.beta_1 ** t)

  @torch.no_grad()
  def apply_gradients(self, params: Dict[str, torch.Tensor],
                      grads: Dict[str, torch.Tensor]) -> None:
    self.iterations += 1
    lr_t = self.step_size(self.iterations)
    for name, w in params.items():
      g = grads[name]
      if name not in self.m:
        self.m[name] = torch.zeros_like(w)
        self.v[name] = torch.zeros_like(w)
      m, v = self.m[name], self.v[name]
      m.mul_(self.beta_1).add_(g, alpha=1.0 - self.beta_1)
      v.mul_(self.beta_2).addcmul_(g, g, value=1.0 - self.beta_2)
      w.sub_(lr_t * m / (v.sqrt() + self.epsilon))

  # ---- checkpoint helpers (keys mirror Keras' optimizer object graph)
  def hyper_state(self) -> Dict[str, float]:
    return {"iter": float(self.iterations), "learning_rate": self.lr,
            "beta_1": self.beta_1, "beta_2": self.beta_2, "decay": 0.0}

  def load_hyper_state(self, st: Dict[str, float]) -> None:
    self.iterations = int(st.get("iter", 0))
    self.lr = float(st.get("learning_rate", self.lr))
    self.beta_1 = float(st.get("beta_1", self.beta_1))
    self.beta_2 = float(st.get("beta_2", self.beta_2))
