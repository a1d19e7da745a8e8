"""Functional PyTorch oracle of the generator / discriminator.

Parameters live in an ordered ``dict[str, Tensor]`` keyed by the Keras
object-graph names from :meth:`NetSpec.variables` (TF layouts). The forward
is a direct interpretation of the :class:`ConvLayer` list — see
cyclegan/model.py:36-126 for the reference's block builders.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Optional

import torch

from ..ops import oracle as O
from .spec import ConvLayer, NetSpec

Params = Dict[str, torch.Tensor]


def init_params(spec: NetSpec,
                generator: Optional[torch.Generator] = None,
                device: str | torch.device = "cpu",
                dtype: torch.dtype = torch.float32) -> Params:
  """Initialise variables like the reference (model.py:10-11, Keras defaults).

  kernels / γ ~ N(0, 0.02); β, biases = 0; the generator's final conv uses
  Keras' default glorot_uniform because no ``kernel_initializer`` is passed
  (model.py:165-166). Sampling always happens on CPU so every device / rank
  gets bit-identical initial weights for a given seed.
  """
  params: Params = OrderedDict()
  for name, shape, init in spec.variables():
    if init == "normal":
      t = torch.empty(shape, dtype=torch.float32).normal_(0.0, 0.02, generator=generator)
    elif init == "zeros":
      t = torch.zeros(shape, dtype=torch.float32)
    elif init == "glorot":
      kh, kw, cin, cout = shape
      limit = math.sqrt(6.0 / (kh * kw * cin + kh * kw * cout))
      t = torch.empty(shape, dtype=torch.float32).uniform_(-limit, limit, generator=generator)
    else:
      raise ValueError(init)
    params[name] = t.to(device=device, dtype=dtype)
  return params


class _Bf16Storage(torch.autograd.Function):
  """A tensor that is STORED in bf16 by the native engine: its value is rounded on the way forward
  and its gradient (also a stored bf16 tensor there) on the way back."""

  @staticmethod
  def forward(ctx, x):
    return x.to(torch.bfloat16).to(x.dtype)

  @staticmethod
  def backward(ctx, g):
    return g.to(torch.bfloat16).to(g.dtype)


def bf16_storage(x: torch.Tensor) -> torch.Tensor:
  return _Bf16Storage.apply(x)


class _MxConv(torch.autograd.Function):
  """Stride-1 "valid" conv whose forward and input-gradient products see MXFP8 operands (E4M3 with one
  power-of-two scale per 32 elements of the contraction axis: input channels forward, OUTPUT channels for
  the gradient wrt the input), the weight gradient full-precision operands — the numerics of the native
  ``--dtype mxfp8`` residual-block convs (csrc/conv_gemm_mx.cu), in plain PyTorch."""

  @staticmethod
  def forward(ctx, xpad, kernel):
    from ..ops import mxfp8
    ctx.save_for_backward(xpad, kernel)
    kq = mxfp8.fake_quantize(kernel.permute(0, 1, 3, 2)).permute(0, 1, 3, 2)     # blocks along cin
    return O.conv2d(mxfp8.fake_quantize(xpad), kq.to(xpad.dtype), 1, "valid")

  @staticmethod
  def backward(ctx, dy):
    from ..ops import mxfp8
    xpad, kernel = ctx.saved_tensors
    dx = dw = None
    with torch.enable_grad():
      if ctx.needs_input_grad[0]:
        xz = torch.zeros_like(xpad, requires_grad=True)
        kt = mxfp8.fake_quantize(kernel).to(xpad.dtype)                             # blocks along cout
        (dx,) = torch.autograd.grad(O.conv2d(xz, kt, 1, "valid"), xz, mxfp8.fake_quantize(dy).to(dy.dtype))
      if ctx.needs_input_grad[1]:
        kk = kernel.detach().requires_grad_(True)
        (dw,) = torch.autograd.grad(O.conv2d(xpad.detach(), kk, 1, "valid"), kk, dy)
    return dx, dw


def mx_layer(l: ConvLayer) -> bool:
  """Layers the native engine runs block-scaled under ``--dtype mxfp8``: the residual-block 3x3 convs."""
  return l.kind == "conv" and l.res in ("a", "b") and l.cin % 128 == 0 and l.cout == 256 and l.pad == "reflect"


def _layer(l: ConvLayer, params: Params, x: torch.Tensor, rnd=None, mx: bool = False) -> torch.Tensor:
  p = f"layer_with_weights-{l.lw_index}"
  kernel = params[f"{p}/kernel"]
  bias = params.get(f"{p}/bias") if l.bias else None
  if l.kind == "conv":
    if mx and mx_layer(l) and bias is None:
      y = _MxConv.apply(O.reflect_pad(x, l.k // 2), kernel)
    elif l.pad == "reflect":
      y = O.conv2d(O.reflect_pad(x, l.k // 2), kernel, l.stride, "valid", bias)
    else:
      y = O.conv2d(x, kernel, l.stride, "same", bias)
  else:
    y = O.conv2d_transpose(x, kernel, l.stride, bias)
  if l.norm:
    if rnd is not None:
      y = rnd(y)     # the raw conv output is a stored tensor; the statistics are those of the stored values
    q = f"layer_with_weights-{l.lw_index + 1}"
    y = O.instance_norm(y, params[f"{q}/gamma"], params[f"{q}/beta"])
  if l.act == "relu":
    y = torch.relu(y)
  elif l.act == "leaky":
    y = O.leaky_relu(y, 0.2)
  elif l.act == "tanh":
    y = torch.tanh(y)
  if rnd is not None and l.res != "b":
    y = rnd(y)
  return y


def forward(spec: NetSpec, params: Params, x: torch.Tensor, rnd=None, mx: bool = False) -> torch.Tensor:
  """NHWC in → NHWC out. ``training`` is irrelevant (no BN / dropout).

  ``rnd``: optional rounding applied wherever the native engine STORES a tensor (raw conv outputs
  of normalised layers, layer outputs, residual sums) — ``bf16_storage`` makes this fp32 oracle see
  the same activations as the bf16 engine, so activation masks agree and what is left of a
  gradient comparison is the backward pass itself. ``mx``: emulate ``--dtype mxfp8`` (block-scaled
  E4M3 operands in the residual-block convs' forward and dgrad products, see ``_MxConv``)."""
  skip = None
  for l in spec.layers:
    if l.res == "a":
      skip = x
    x = _layer(l, params, x, rnd, mx)
    if l.res == "b":
      x = skip + x  # layers.add([inputs, outputs]); no activation after (model.py:73)
      if rnd is not None:
        x = rnd(x)
  return x
