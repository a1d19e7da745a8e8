"""Architecture parity: param counts, variable counts, shapes, FLOPs."""
import torch

from tf2_cyclegan_b200.models import oracle_net as N
from tf2_cyclegan_b200.models.spec import discriminator_spec, generator_spec


def test_param_and_variable_counts():
  g, d = generator_spec(), discriminator_spec()
  assert g.num_params() == 11_383_427 and len(g.variables()) == 71
  assert d.num_params() == 2_765_633 and len(d.variables()) == 13
  assert generator_spec(num_residual_blocks=1).num_params() == 1_938_051


def test_layer_with_weights_numbering():
  g, d = generator_spec(), discriminator_spec()
  assert g.layers[-1].lw_index == 46 and g.layers[0].lw_index == 0
  assert [l.lw_index for l in d.layers] == [0, 1, 3, 5, 7]
  names = [n for n, _, _ in d.variables()]
  assert names[0] == "layer_with_weights-0/kernel" and names[1] == "layer_with_weights-0/bias"
  assert "layer_with_weights-2/gamma" in names and "layer_with_weights-7/bias" in names


def test_flops():
  assert abs(generator_spec().forward_flops(256, 256) / 1e9 - 99.1) < 0.1
  assert abs(discriminator_spec().forward_flops(256, 256) / 1e9 - 6.56) < 0.01


def test_output_shapes_and_init():
  gen = torch.Generator().manual_seed(0)
  g = generator_spec(num_residual_blocks=1)
  p = N.init_params(g, gen)
  y = N.forward(g, p, torch.randn(2, 32, 32, 3))
  assert y.shape == (2, 32, 32, 3) and y.abs().max() <= 1.0
  d = discriminator_spec()
  pd = N.init_params(d, gen)
  assert N.forward(d, pd, torch.randn(1, 256, 256, 3)).shape == (1, 32, 32, 1)
  # gamma ~ N(0, 0.02) — around 0, not 1 (model.py:10-11); beta and biases zero
  gam = p["layer_with_weights-1/gamma"]
  assert gam.abs().mean() < 0.05 and p["layer_with_weights-1/beta"].abs().sum() == 0
  # final conv: glorot_uniform + zero bias
  k = p[f"layer_with_weights-{g.layers[-1].lw_index}/kernel"]
  limit = (6.0 / (49 * 64 + 49 * 3)) ** 0.5
  assert k.abs().max() <= limit and k.abs().max() > 0.5 * limit


def test_residual_has_no_post_add_activation():
  g = generator_spec(num_residual_blocks=1)
  b = [l for l in g.layers if l.res == "b"][0]
  assert b.act is None and b.norm
