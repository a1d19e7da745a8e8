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


def _layer(l: ConvLayer, params: Params, x: torch.Tensor) -> torch.Tensor:
  p = f"layer_with_weights-{l.lw_index}"
  kernel = params[f"{p}/kernel"]
  bias = params.get(f"{p}/bias") if l.bias else None
  if l.kind == "conv":
    if l.pad == "reflect":
      y = O.conv2d(O.reflect_pad(x, l.k // 2), kernel, l.stride, "valid", bias)
    else:
      y = O.conv2d(x, kernel, l.stride, "same", bias)
  else:
    y = O.conv2d_transpose(x, kernel, l.stride, bias)
  if l.norm:
    q = f"layer_with_weights-{l.lw_index + 1}"
    y = O.instance_norm(y, params[f"{q}/gamma"], params[f"{q}/beta"])
  if l.act == "relu":
    y = torch.relu(y)
  elif l.act == "leaky":
    y = O.leaky_relu(y, 0.2)
  elif l.act == "tanh":
    y = torch.tanh(y)
  return y


def forward(spec: NetSpec, params: Params, x: torch.Tensor) -> torch.Tensor:
  """NHWC in → NHWC out. ``training`` is irrelevant (no BN / dropout)."""
  skip = None
  for l in spec.layers:
    if l.res == "a":
      skip = x
    x = _layer(l, params, x)
    if l.res == "b":
      x = skip + x  # layers.add([inputs, outputs]); no activation after (model.py:73)
  return x
