"""The oracle's MXFP8 emulation (models/oracle_net._MxConv, OracleBackend(emulate_mxfp8=True)): the CPU-runnable
model of what ``--dtype mxfp8`` computes on the GPU."""
import types

import torch

from tf2_cyclegan_b200.models import oracle_net
from tf2_cyclegan_b200.ops import mxfp8
from tf2_cyclegan_b200.ops import oracle as O


def test_mx_conv_is_exact_on_representable_operands_and_quantises_otherwise():
  torch.manual_seed(0)
  # values with <= 3 mantissa bits and one magnitude class per 32-block survive the round trip unchanged
  x = torch.randint(-7, 8, (1, 6, 6, 64)).float().requires_grad_(True)
  k = (torch.randint(-7, 8, (3, 3, 64, 32)).float() / 8).requires_grad_(True)
  dy = torch.randint(-7, 8, (1, 4, 4, 32)).float()
  assert torch.equal(mxfp8.fake_quantize(x.detach()), x.detach())
  y = oracle_net._MxConv.apply(x, k)
  ref = O.conv2d(x, k, 1, "valid")
  assert torch.equal(y, ref)
  gx, gk = torch.autograd.grad(y, (x, k), dy)
  rx, rk = torch.autograd.grad(ref, (x, k), dy)
  assert torch.allclose(gx, rx) and torch.allclose(gk, rk)
  # generic operands: close to, but not equal to, the unquantised product; the weight gradient is NOT quantised
  x2 = torch.randn(1, 6, 6, 64, requires_grad=True)
  k2 = (torch.randn(3, 3, 64, 32) * 0.1).requires_grad_(True)
  dy2 = torch.randn(1, 4, 4, 32)
  y2, r2 = oracle_net._MxConv.apply(x2, k2), O.conv2d(x2, k2, 1, "valid")
  e = ((y2 - r2).norm() / r2.norm()).item()
  assert 1e-4 < e < 5e-2
  gx2, gk2 = torch.autograd.grad(y2, (x2, k2), dy2)
  rx2, rk2 = torch.autograd.grad(r2, (x2, k2), dy2)
  assert 1e-4 < ((gx2 - rx2).norm() / rx2.norm()).item() < 5e-2
  assert torch.allclose(gk2, rk2, rtol=1e-5, atol=1e-6)


def test_oracle_backend_emulates_mxfp8_on_cpu():
  from tf2_cyclegan_b200.engine.cyclegan import build_specs
  from tf2_cyclegan_b200.engine.oracle_backend import OracleBackend
  from tf2_cyclegan_b200.parallel.strategy import Strategy
  st = Strategy(device="cpu")
  specs = build_specs(types.SimpleNamespace(num_residual_blocks=1))
  assert sum(oracle_net.mx_layer(l) for l in specs["G"].layers) == 2 and not any(map(oracle_net.mx_layer, specs["X"].layers))
  a = OracleBackend(specs, 1, st, seed=3)
  b = OracleBackend(specs, 1, st, seed=3, emulate_mxfp8=True)
  torch.manual_seed(1)
  x = torch.rand(1, 32, 32, 3) * 2 - 1
  y = torch.rand(1, 32, 32, 3) * 2 - 1
  ra, rb = a.test_step(x, y), b.test_step(x, y)
  assert any(ra[k].item() != rb[k].item() for k in ra)            # the emulation is active ...
  for k in ra:                                                   # ... and a small perturbation
    assert abs(ra[k].item() - rb[k].item()) < 5e-2 * max(abs(ra[k].item()), 0.05), (k, ra[k], rb[k])
  _, ga = a.local_gradients(x, y)
  _, gb = b.local_gradients(x, y)
  for n in ("G", "F"):
    err = ((ga[n] - gb[n]).norm() / ga[n].norm()).item()
    assert 0 < err < 0.2, (n, err)


def test_cli_accepts_mxfp8_for_the_oracle_but_not_fp8():
  import pytest
  from tf2_cyclegan_b200.cli import parse_args
  a = parse_args(["--output_dir", "/tmp/x", "--backend", "oracle", "--dtype", "mxfp8"])
  assert a.dtype == "mxfp8"
  with pytest.raises(SystemExit):
    parse_args(["--output_dir", "/tmp/x", "--backend", "oracle", "--dtype", "fp8"])


def test_facade_builds_the_emulating_oracle(tmp_path):
  from tf2_cyclegan_b200.engine.cyclegan import CycleGAN
  from tf2_cyclegan_b200.parallel.strategy import Strategy
  args = types.SimpleNamespace(output_dir=str(tmp_path), global_batch_size=1, batch_size=1, image_size=32,
                               num_residual_blocks=1, backend="oracle", dtype="mxfp8", seed=1234)
  gan = CycleGAN(args, Strategy(device="cpu"))
  assert gan.backend.name == "oracle" and gan.backend.mx and gan.backend.compute_dtype == torch.float32
