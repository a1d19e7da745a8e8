"""Gradient routing of the reference's train_step."""
import torch

from tf2_cyclegan_b200.engine.cyclegan import build_specs
from tf2_cyclegan_b200.engine.oracle_backend import NETS, OracleBackend
from tf2_cyclegan_b200.parallel.strategy import Strategy


class A:
  num_residual_blocks = 1


def make(gbs=2):
  return OracleBackend(build_specs(A()), gbs, Strategy(device="cpu"), seed=7)


def test_each_generator_only_gets_its_own_total():
  be = make()
  torch.manual_seed(0)
  x, y = torch.randn(2, 16, 16, 3), torch.randn(2, 16, 16, 3)
  res, grads = be.local_gradients(x, y)
  # independent re-derivation: dF_total/dG must be ignored, so G's gradient equals
  # the gradient of (G_loss + G_cycle + G_identity) with the inner F(y) treated as data
  L = be.losses
  fake_x = be.net("F", y).detach()
  G_total = (L.generator_loss(be.net("Y", be.net("G", x))) +
             L.cycle_loss(y, be.net("G", fake_x)) + L.identity_loss(y, be.net("G", y)))
  g = torch.autograd.grad(G_total, list(be.params["G"].values()))
  ref = torch.cat([t.reshape(-1) for t in g])
  assert torch.allclose(grads["G"], ref, rtol=1e-4, atol=1e-7)
  assert abs(res["loss_G/total"].item() - G_total.item()) < 1e-5


def test_discriminator_grads_ignore_generators_but_use_undetached_fakes():
  be = make()
  torch.manual_seed(1)
  x, y = torch.randn(2, 16, 16, 3), torch.randn(2, 16, 16, 3)
  _, grads = be.local_gradients(x, y)
  L = be.losses
  fake_x = be.net("F", y).detach()
  X_loss = L.discriminator_loss(be.net("X", x), be.net("X", fake_x))
  g = torch.autograd.grad(X_loss, list(be.params["X"].values()))
  assert torch.allclose(grads["X"], torch.cat([t.reshape(-1) for t in g]), rtol=1e-4, atol=1e-7)


def test_updates_use_pre_update_weights_and_step_changes_all_nets():
  be = make()
  before = {n: torch.cat([p.detach().reshape(-1).clone() for p in be.params[n].values()]) for n in NETS}
  x, y = torch.randn(2, 16, 16, 3), torch.randn(2, 16, 16, 3)
  res = be.train_step(x, y)
  assert len(res) == 10
  for n in NETS:
    after = torch.cat([p.detach().reshape(-1) for p in be.params[n].values()])
    assert not torch.equal(before[n], after)
    assert be.opt[n].iterations == 1


def test_test_step_has_14_scalars():
  be = make()
  r = be.test_step(torch.randn(1, 16, 16, 3), torch.randn(1, 16, 16, 3))
  assert len(r) == 14 and "error/MAE(X, F(G(X)))" in r
