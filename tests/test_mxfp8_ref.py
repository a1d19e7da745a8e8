"""CPU checks of the MXFP8 reference quantiser (ops/mxfp8.py) that the GPU numerics tests compare the
CUDA writers against: scale rule, round-trip error, weight-scale fetch order."""
import torch

from tf2_cyclegan_b200.ops import mxfp8


def test_scale_rule_and_roundtrip():
  torch.manual_seed(0)
  t = torch.randn(64, 256) * torch.exp2(torch.randint(-12, 12, (64, 8)).float()).repeat_interleave(32, dim=1)
  q, sb = mxfp8.quantize(t)
  assert q.dtype == torch.uint8 and q.shape == t.shape and sb.shape == (64, 8)
  s = mxfp8.scale_values(sb)
  amax = t.reshape(64, 8, 32).abs().amax(-1)
  assert (amax / s <= mxfp8.E4M3_MAX).all()            # nothing saturates
  assert (amax / s > mxfp8.E4M3_MAX / 2).all()         # and the scale is the smallest such power of two
  back = mxfp8.dequantize(q, sb)
  # E4M3 has 3 mantissa bits: relative rounding error <= 2^-4 for normal values; elements far below the
  # block maximum fall into the subnormal range (absolute error <= 2^-10 of the scaled range)
  err = (back - t).abs()
  bound = torch.maximum(t.abs() * 2.0 ** -4, (s * 2.0 ** -10).repeat_interleave(32, dim=1))
  assert (err <= bound * 1.0001).all()


def test_edge_values():
  z = torch.zeros(1, 32)
  q, sb = mxfp8.quantize(z)
  assert sb.item() == 0 and (q == 0).all() and (mxfp8.dequantize(q, sb) == 0).all()
  one = torch.zeros(1, 32); one[0, 3] = 448.0
  q, sb = mxfp8.quantize(one)
  assert sb.item() == 127 and mxfp8.dequantize(q, sb)[0, 3].item() == 448.0     # scale 1, exact
  over = torch.zeros(1, 32); over[0, 0] = 449.0
  q, sb = mxfp8.quantize(over)
  assert sb.item() == 128                                                       # next power of two


def test_weight_scale_order():
  cout, taps, cin = 256, 9, 256
  sb = torch.arange(cout * taps * cin // 32, dtype=torch.int64).reshape(cout, taps * cin // 32)
  flat = mxfp8.weight_scale_order(sb, taps, cin)
  for n, tap, cb, k in [(0, 0, 0, 0), (37, 4, 1, 2), (200, 8, 0, 3), (255, 8, 1, 3)]:
    chunk = ((n // 128) * (cin // 128) + cb) * taps + tap
    pos = chunk * 512 + (n % 32) * 16 + ((n % 128) // 32) * 4 + k
    assert flat[pos].item() == sb[n, tap * (cin // 32) + cb * 4 + k].item()
