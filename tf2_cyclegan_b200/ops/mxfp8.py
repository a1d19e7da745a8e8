"""MXFP8 (OCP microscaling) reference in plain PyTorch: E4M3 elements with one E8M0 power-of-two scale
per 32 consecutive elements of the last dimension. Bit-exact model of what the CUDA side writes
(``csrc/mx.cuh``: ``mx_scale_byte`` / ``mx_inv_scale``; ``in_apply_kernel<.., 2>``;
``mx_weight_quant_kernel``) — used by the numerics tests and to build oracles of the block-scaled
convolution (``csrc/conv_gemm_mx.cu``, BASELINE config 3).
"""
from __future__ import annotations

from typing import Tuple

import torch

BLOCK = 32          # elements per scale
E4M3_MAX = 448.0
SF_RECORD = 16      # bytes per pixel in an activation scale buffer


def scale_bytes(amax: torch.Tensor) -> torch.Tensor:
  """E8M0 byte (value 2^(byte - 127)) of the smallest power of two s with amax / s <= 448."""
  b = amax.float().contiguous().view(torch.int32)
  e = (b >> 23) - 8 + ((b & 0x7FFFFF) > 0x600000).to(torch.int32)
  return e.clamp(0, 254).to(torch.uint8)


def scale_values(sb: torch.Tensor) -> torch.Tensor:
  return torch.ldexp(torch.ones_like(sb, dtype=torch.float32), sb.to(torch.int32) - 127)


def quantize(t: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
  """t [..., K] (K % 32 == 0) -> (E4M3 bytes [..., K] uint8, scale bytes [..., K / 32] uint8)."""
  k = t.shape[-1]
  assert k % BLOCK == 0
  blocks = t.float().reshape(*t.shape[:-1], k // BLOCK, BLOCK)
  sb = scale_bytes(blocks.abs().amax(-1))
  inv = torch.ldexp(torch.ones_like(sb, dtype=torch.float32), 127 - sb.to(torch.int32))
  q = (blocks * inv[..., None]).to(torch.float8_e4m3fn)
  return q.view(torch.uint8).reshape(t.shape), sb


def dequantize(q: torch.Tensor, sb: torch.Tensor) -> torch.Tensor:
  k = q.shape[-1]
  v = q.view(torch.float8_e4m3fn).float().reshape(*q.shape[:-1], k // BLOCK, BLOCK)
  return (v * scale_values(sb)[..., None]).reshape(q.shape)


def fake_quantize(t: torch.Tensor) -> torch.Tensor:
  """Round-trip through MXFP8 (what the block-scaled MMA sees), as fp32."""
  return dequantize(*quantize(t))


def weight_scale_order(sb: torch.Tensor, taps: int, cin: int) -> torch.Tensor:
  """Scale bytes [cout, taps * cin / 32] of a [cout, taps * cin] weight matrix -> the flat order
  the conv kernels fetch: [cout / 128][cin / 128][tap][cout % 32][(cout % 128) / 32][K slice of 4]
  (all taps of one 128-channel block are contiguous: one bulk copy per block)."""
  cout = sb.shape[0]
  assert cout % 128 == 0 and cin % 128 == 0 and sb.shape[1] == taps * cin // BLOCK
  v = sb.reshape(cout // 128, 4, 32, taps, cin // 128, 4)      # [nt, n1, n0, tap, cb, k]
  return v.permute(0, 4, 3, 2, 1, 5).contiguous().reshape(-1)  # [nt, cb, tap, n0, n1, k]
