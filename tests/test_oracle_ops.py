"""Op semantics vs closed forms / tiny golden values (SURVEY.md §4 'unit (CPU)')."""
import math

import torch

from tf2_cyclegan_b200.ops import oracle as O
from tf2_cyclegan_b200.ops.keras_adam import KerasAdam
from tf2_cyclegan_b200.ops.losses import LossHeads


def test_same_pad_is_asymmetric():
  assert O.same_pad_amounts(256, 3, 2) == (0, 1)     # generator downsample
  assert O.same_pad_amounts(256, 4, 2) == (1, 1)     # PatchGAN s2
  assert O.same_pad_amounts(32, 4, 1) == (1, 2)      # PatchGAN s1
  assert O.same_pad_amounts(255, 3, 2) == (1, 1)


def test_conv_same_matches_manual():
  torch.manual_seed(0)
  x = torch.randn(1, 6, 6, 2)
  k = torch.randn(3, 3, 2, 4)
  y = O.conv2d(x, k, stride=2, padding="same")
  assert y.shape == (1, 3, 3, 4)
  xp = torch.zeros(1, 7, 7, 2)
  xp[:, :6, :6] = x   # pad (0, 1)
  ref = torch.zeros(1, 3, 3, 4)
  for oh in range(3):
    for ow in range(3):
      patch = xp[0, 2 * oh:2 * oh + 3, 2 * ow:2 * ow + 3]  # [3,3,2]
      ref[0, oh, ow] = torch.einsum("hwc,hwco->o", patch, k)
  assert torch.allclose(y, ref, atol=1e-5)


def test_discriminator_keeps_32x32():
  x = torch.randn(1, 32, 32, 8)
  y = O.conv2d(x, torch.randn(4, 4, 8, 1), stride=1, padding="same")
  assert y.shape == (1, 32, 32, 1)


def test_conv_transpose_is_adjoint_of_same_conv():
  torch.manual_seed(1)
  k = torch.randn(3, 3, 5, 7)               # HWIO for the conv: in=5, out=7
  x = torch.randn(2, 8, 8, 5)
  y = torch.randn(2, 4, 4, 7)
  fwd = O.conv2d(x, k, stride=2, padding="same")
  # Conv2DTranspose kernel is [kh, kw, out, in] with out=5 (conv's in), in=7
  bwd = O.conv2d_transpose(y, k, stride=2)
  assert bwd.shape == x.shape
  assert torch.allclose((fwd * y).sum(), (x * bwd).sum(), rtol=1e-4)


def test_reflect_pad():
  x = torch.arange(9.0).reshape(1, 3, 3, 1)
  y = O.reflect_pad(x, 1)[0, :, :, 0]
  assert y.shape == (5, 5)
  assert y[0].tolist() == [4, 3, 4, 5, 4]
  assert y[:, 0].tolist() == [4, 1, 4, 7, 4]


def test_instance_norm_eps_and_biased_var():
  torch.manual_seed(2)
  x = torch.randn(2, 5, 7, 3) * 3 + 1
  g = torch.randn(3)
  b = torch.randn(3)
  y = O.instance_norm(x, g, b)
  mean = x.mean((1, 2), keepdim=True)
  var = ((x - mean) ** 2).mean((1, 2), keepdim=True)
  ref = (x - mean) / torch.sqrt(var + 1e-3) * g + b
  assert torch.allclose(y, ref, atol=1e-5)
  assert O.IN_EPS == 1e-3


def test_losses_divide_by_global_batch():
  L = LossHeads(global_batch_size=8)
  d = torch.full((2, 4, 4, 1), 0.5)
  assert math.isclose(L.generator_loss(d).item(), 2 * 0.25 / 8, rel_tol=1e-6)
  x = torch.zeros(2, 4, 4, 3)
  y = torch.ones(2, 4, 4, 3)
  assert math.isclose(L.cycle_loss(x, y).item(), 10 * 2 / 8, rel_tol=1e-6)
  assert math.isclose(L.identity_loss(x, y).item(), 5 * 2 / 8, rel_tol=1e-6)
  assert math.isclose(L.discriminator_loss(d, d).item(), 0.5 * (0.25 + 0.25) * 2 / 8, rel_tol=1e-6)


def test_bce_matches_torch():
  p = torch.rand(3, 5).clamp(0.01, 0.99)
  t = (torch.rand(3, 5) > 0.5).float()
  ref = torch.nn.functional.binary_cross_entropy(p, t, reduction="none").mean(1)
  assert torch.allclose(O.bce(t, p), ref, atol=1e-6)


def test_keras_adam_formula():
  opt = KerasAdam()
  assert (opt.lr, opt.beta_1, opt.beta_2, opt.epsilon) == (2e-4, 0.5, 0.9, 1e-7)
  w = {"w": torch.tensor([1.0, -2.0])}
  g = {"w": torch.tensor([0.5, -0.25])}
  w0 = w["w"].clone()
  m = torch.zeros(2)
  v = torch.zeros(2)
  for t in range(1, 4):
    opt.apply_gradients(w, g)
    m = 0.5 * m + 0.5 * g["w"]
    v = 0.9 * v + 0.1 * g["w"] ** 2
    lr_t = 2e-4 * math.sqrt(1 - 0.9 ** t) / (1 - 0.5 ** t)
    w0 = w0 - lr_t * m / (v.sqrt() + 1e-7)
    assert torch.allclose(w["w"], w0, atol=1e-9)
  assert opt.iterations == 3
