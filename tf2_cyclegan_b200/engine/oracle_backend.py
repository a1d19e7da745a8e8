"""Autograd (PyTorch eager) backend: the numerical oracle of the whole step.

Implements exactly the reference's ``train_step`` / ``test_step`` /
``cycle_step`` (main.py:197-323) including its gradient routing: each of the
four optimizers differentiates its own loss w.r.t. its own network's
variables only, all from pre-update weights (one persistent tape,
main.py:209,249-260). Runs on CPU (plumbing config, tests) and on CUDA through
ATen/cuDNN — which is also what the stand-in baseline in ``baseline/`` measures.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict

import torch

from ..models import oracle_net
from ..models.spec import NetSpec
from ..ops.keras_adam import KerasAdam
from ..ops.losses import LossHeads
from ..ops import oracle as O
from ..parallel.strategy import Strategy

NETS = ("G", "F", "X", "Y")


def flatten(params: Dict[str, torch.Tensor], grads) -> torch.Tensor:
  return torch.cat([g.reshape(-1) for g in grads]) if grads else torch.zeros(0)


class OracleBackend:
  name = "oracle"

  def __init__(self, specs: Dict[str, NetSpec], global_batch_size: int,
               strategy: Strategy, seed: int = 1234,
               compute_dtype: torch.dtype = torch.float32, emulate_bf16_storage: bool = False,
               emulate_mxfp8: bool = False):
    self.specs = specs
    self.strategy = strategy
    self.device = strategy.device
    self.losses = LossHeads(global_batch_size)
    self.compute_dtype = compute_dtype
    # fp32 math on activations rounded to bf16 wherever the native engine stores them (tests)
    self.rnd = oracle_net.bf16_storage if emulate_bf16_storage else None
    # the numerics of the native ``--dtype mxfp8``: block-scaled E4M3 operands in the generators' residual convs
    self.mx = bool(emulate_mxfp8)
    gen = torch.Generator().manual_seed(seed)
    self.params: Dict[str, Dict[str, torch.Tensor]] = OrderedDict()
    self.opt: Dict[str, KerasAdam] = OrderedDict()
    for n in NETS:
      p = oracle_net.init_params(specs[n], generator=gen, device=self.device)
      for t in p.values():
        t.requires_grad_(True)
      self.params[n] = p
      self.opt[n] = KerasAdam()

  # ------------------------------------------------------------- networks
  def net(self, n: str, x: torch.Tensor) -> torch.Tensor:
    if self.compute_dtype != torch.float32 and self.device.type == "cuda":
      with torch.autocast("cuda", dtype=self.compute_dtype):
        return oracle_net.forward(self.specs[n], self.params[n], x, mx=self.mx).float()
    if self.rnd is not None:
      x = self.rnd(x)                                  # network inputs are stored as bf16 pixels
    return oracle_net.forward(self.specs[n], self.params[n], x, self.rnd, self.mx)

  @torch.no_grad()
  def cycle_step(self, x: torch.Tensor, y: torch.Tensor):
    """main.py:197-205 — returns (fake_x, fake_y, cycle_x, cycle_y)."""
    fake_y = self.net("G", x)
    cycle_x = self.net("F", fake_y)
    fake_x = self.net("F", y)
    cycle_y = self.net("G", fake_x)
    return fake_x, fake_y, cycle_x, cycle_y

  # ---------------------------------------------------------------- losses
  def _forward_losses(self, x, y, with_errors: bool):
    L = self.losses
    fake_y = self.net("G", x)
    fake_x = self.net("F", y)
    d_fake_x = self.net("X", fake_x)   # computed once; the reference computes it
    d_fake_y = self.net("Y", fake_y)   # twice with identical inputs (main.py:213,241)
    G_loss = L.generator_loss(d_fake_y)
    F_loss = L.generator_loss(d_fake_x)
    cycle_y = self.net("G", fake_x)
    cycle_x = self.net("F", fake_y)
    G_cycle = L.cycle_loss(y, cycle_y)
    F_cycle = L.cycle_loss(x, cycle_x)
    same_y = self.net("G", y)
    same_x = self.net("F", x)
    G_identity = L.identity_loss(y, same_y)
    F_identity = L.identity_loss(x, same_x)
    G_total = G_loss + G_cycle + G_identity
    F_total = F_loss + F_cycle + F_identity
    X_loss = L.discriminator_loss(self.net("X", x), d_fake_x)
    Y_loss = L.discriminator_loss(self.net("Y", y), d_fake_y)
    result = OrderedDict([
        ("loss_G/loss", G_loss), ("loss_G/cycle", G_cycle),
        ("loss_G/identity", G_identity), ("loss_G/total", G_total),
        ("loss_F/loss", F_loss), ("loss_F/cycle", F_cycle),
        ("loss_F/identity", F_identity), ("loss_F/total", F_total),
        ("loss_X/loss", X_loss), ("loss_Y/loss", Y_loss)])
    if with_errors:  # main.py:317-320
      result["error/MAE(X, F(G(X)))"] = L.reduce_mean(O.mae(x, cycle_x))
      result["error/MAE(Y, G(F(Y)))"] = L.reduce_mean(O.mae(y, cycle_y))
      result["error/MAE(X, F(X))"] = L.reduce_mean(O.mae(x, same_x))
      result["error/MAE(Y, G(Y))"] = L.reduce_mean(O.mae(y, same_y))
    return result

  def local_gradients(self, x: torch.Tensor, y: torch.Tensor):
    """Per-replica losses and flat per-network gradients (no collectives)."""
    x = x.to(self.device, torch.float32)
    y = y.to(self.device, torch.float32)
    result = self._forward_losses(x, y, with_errors=False)
    targets = {"G": result["loss_G/total"], "F": result["loss_F/total"],
               "X": result["loss_X/loss"], "Y": result["loss_Y/loss"]}
    grads = {}
    for i, n in enumerate(NETS):
      plist = list(self.params[n].values())
      if x.shape[0] == 0:
        g = [torch.zeros_like(p) for p in plist]
      else:
        g = torch.autograd.grad(targets[n], plist, retain_graph=(i < len(NETS) - 1))
      grads[n] = torch.cat([t.reshape(-1) for t in g])
    return {k: v.detach() for k, v in result.items()}, grads

  def apply_flat_gradients(self, n: str, flat: torch.Tensor) -> None:
    grads, off = {}, 0
    for name, p in self.params[n].items():
      grads[name] = flat[off:off + p.numel()].view_as(p)
      off += p.numel()
    self.opt[n].apply_gradients({k: v.data for k, v in self.params[n].items()}, grads)

  def train_step(self, x: torch.Tensor, y: torch.Tensor) -> Dict[str, torch.Tensor]:
    """main.py:207-262 + the cross-replica SUM of ``minimize``."""
    result, grads = self.local_gradients(x, y)
    for n in NETS:  # one packed all-reduce per network, like NcclAllReduce(num_packs=1)
      self.strategy.all_reduce_sum(grads[n])
      self.apply_flat_gradients(n, grads[n])
    return result

  @torch.no_grad()
  def test_step(self, x: torch.Tensor, y: torch.Tensor) -> Dict[str, torch.Tensor]:
    x = x.to(self.device, torch.float32)
    y = y.to(self.device, torch.float32)
    return self._forward_losses(x, y, with_errors=True)

  # ------------------------------------------------------------ checkpoint
  def state_dict(self) -> Dict[str, torch.Tensor]:
    """Logical (TF-layout) tensors keyed like the Keras object graph."""
    sd: Dict[str, torch.Tensor] = OrderedDict()
    for n in NETS:
      for name, p in self.params[n].items():
        sd[f"{n}/{name}/.ATTRIBUTES/VARIABLE_VALUE"] = p.detach().cpu()
      opt = self.opt[n]
      for hk, hv in opt.hyper_state().items():
        sd[f"{n}_optimizer/{hk}/.ATTRIBUTES/VARIABLE_VALUE"] = torch.tensor(hv, dtype=torch.float64)
      for name, p in self.params[n].items():
        for slot, store in (("m", opt.m), ("v", opt.v)):
          val = store.get(name)
          val = torch.zeros_like(p) if val is None else val
          sd[f"{n}/{name}/.OPTIMIZER_SLOT/{n}_optimizer/{slot}/.ATTRIBUTES/VARIABLE_VALUE"] = val.detach().cpu()
    return sd

  def load_state_dict(self, sd: Dict[str, torch.Tensor], expect_partial: bool = False) -> None:
    """Like ``checkpoint.read()`` (main.py:162-170): missing optimizer state is tolerated (warned
    about, starts from zero); missing weights raise unless ``expect_partial``."""
    missing, missing_opt = [], []
    for n in NETS:
      for name, p in self.params[n].items():
        key = f"{n}/{name}/.ATTRIBUTES/VARIABLE_VALUE"
        if key in sd:
          p.data.copy_(sd[key].to(p.device, p.dtype))
        else:
          missing.append(key)
      opt = self.opt[n]
      hyper = {}
      for hk in ("iter", "learning_rate", "beta_1", "beta_2", "decay"):
        key = f"{n}_optimizer/{hk}/.ATTRIBUTES/VARIABLE_VALUE"
        if key in sd:
          hyper[hk] = float(sd[key])
        else:
          missing_opt.append(key)
      opt.load_hyper_state(hyper)
      for name, p in self.params[n].items():
        for slot, store in (("m", opt.m), ("v", opt.v)):
          key = f"{n}/{name}/.OPTIMIZER_SLOT/{n}_optimizer/{slot}/.ATTRIBUTES/VARIABLE_VALUE"
          if key in sd:
            store[name] = sd[key].to(p.device, p.dtype).clone()
          else:
            missing_opt.append(key)
    if missing_opt and self.strategy.rank == 0:
      print(f"[checkpoint] {len(missing_opt)} optimizer entries missing (e.g. {missing_opt[0]}): "
            "Adam state for them starts from zero")
    if missing and not expect_partial:
      raise KeyError(f"checkpoint is missing {len(missing)} weight tensors, e.g. {missing[:3]} "
                     f"(pass expect_partial=True to ignore)")
