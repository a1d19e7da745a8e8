"""Architecture descriptors for the CycleGAN generator and discriminator.

A network is a flat list of :class:`ConvLayer` records; both the PyTorch
oracle (models/oracle_net.py) and the sm_100a planner
(engine/native_plan.py) are driven by the same list, so the two back-ends
cannot drift apart.

Parity with the reference builders:
  * ``generator_spec``      ↔ cyclegan/model.py:129-169 (get_generator)
  * ``discriminator_spec``  ↔ cyclegan/model.py:172-213 (get_discriminator)
  * variable order / names  ↔ Keras object-graph ``layer_with_weights-<i>``
    numbering (conv, IN, conv, IN, …) so checkpoints keep the same logical keys.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple


@dataclass(frozen=True)
class ConvLayer:
  name: str
  kind: str                 # "conv" | "convT"
  k: int
  stride: int
  cin: int
  cout: int
  pad: str                  # "reflect" (pad k//2 then valid) | "same" (TF SAME)
  bias: bool
  norm: bool                # InstanceNormalization after the conv
  act: Optional[str]        # "relu" | "leaky" | "tanh" | None
  kernel_init: str          # "normal" (N(0,0.02)) | "glorot" (Keras default)
  res: Optional[str] = None  # "a": first conv of a residual block, "b": second (+skip)
  lw_index: int = -1        # layer_with_weights index of the conv (IN is +1)

  @property
  def kernel_shape(self) -> Tuple[int, int, int, int]:
    if self.kind == "conv":
      return (self.k, self.k, self.cin, self.cout)      # HWIO
    return (self.k, self.k, self.cout, self.cin)        # Conv2DTranspose: HW,out,in

  def out_hw(self, h: int, w: int) -> Tuple[int, int]:
    if self.kind == "convT":
      return h * self.stride, w * self.stride
    if self.pad == "reflect":
      return h, w
    return -(-h // self.stride), -(-w // self.stride)


@dataclass(frozen=True)
class NetSpec:
  name: str
  kind: str                 # "generator" | "discriminator"
  layers: Tuple[ConvLayer, ...]
  in_channels: int = 3

  def variables(self) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(name, shape, init) for every trainable variable, in Keras order."""
    out = []
    for l in self.layers:
      p = f"layer_with_weights-{l.lw_index}"
      out.append((f"{p}/kernel", l.kernel_shape, l.kernel_init))
      if l.bias:
        out.append((f"{p}/bias", (l.cout,), "zeros"))
      if l.norm:
        q = f"layer_with_weights-{l.lw_index + 1}"
        out.append((f"{q}/gamma", (l.cout,), "normal"))
        out.append((f"{q}/beta", (l.cout,), "zeros"))
    return out

  def num_params(self) -> int:
    n = 0
    for _, shape, _ in self.variables():
      m = 1
      for s in shape:
        m *= s
      n += m
    return n

  def forward_flops(self, h: int, w: int) -> float:
    """Multiply-add FLOPs (2·MACs) of one forward pass on an h×w image."""
    total = 0.0
    for l in self.layers:
      oh, ow = l.out_hw(h, w)
      if l.kind == "conv":
        total += 2.0 * oh * ow * l.k * l.k * l.cin * l.cout
      else:  # every input pixel touches k*k taps
        total += 2.0 * h * w * l.k * l.k * l.cin * l.cout
      h, w = oh, ow
    return total


def generator_spec(filters: int = 64,
                   num_downsampling_blocks: int = 2,
                   num_residual_blocks: int = 9,
                   num_upsample_blocks: int = 2,
                   name: str = "generator") -> NetSpec:
  layers: List[ConvLayer] = []
  idx = 0

  def add(**kw):
    nonlocal idx
    layer = ConvLayer(lw_index=idx, **kw)
    layers.append(layer)
    idx += 2 if layer.norm else 1

  f = filters
  add(name="stem", kind="conv", k=7, stride=1, cin=3, cout=f, pad="reflect",
      bias=False, norm=True, act="relu", kernel_init="normal")
  for i in range(num_downsampling_blocks):
    add(name=f"down{i}", kind="conv", k=3, stride=2, cin=f, cout=2 * f,
        pad="same", bias=False, norm=True, act="relu", kernel_init="normal")
    f *= 2
  for i in range(num_residual_blocks):
    add(name=f"res{i}a", kind="conv", k=3, stride=1, cin=f, cout=f,
        pad="reflect", bias=False, norm=True, act="relu", kernel_init="normal",
        res="a")
    add(name=f"res{i}b", kind="conv", k=3, stride=1, cin=f, cout=f,
        pad="reflect", bias=False, norm=True, act=None, kernel_init="normal",
        res="b")
  for i in range(num_upsample_blocks):
    add(name=f"up{i}", kind="convT", k=3, stride=2, cin=f, cout=f // 2,
        pad="same", bias=False, norm=True, act="relu", kernel_init="normal")
    f //= 2
  # final block: default Keras init (glorot_uniform) and a bias (model.py:164-167)
  add(name="final", kind="conv", k=7, stride=1, cin=f, cout=3, pad="reflect",
      bias=True, norm=False, act="tanh", kernel_init="glorot")
  return NetSpec(name=name, kind="generator", layers=tuple(layers))


def discriminator_spec(filters: int = 64,
                       num_downsampling: int = 3,
                       name: str = "discriminator") -> NetSpec:
  layers: List[ConvLayer] = []
  idx = 0

  def add(**kw):
    nonlocal idx
    layer = ConvLayer(lw_index=idx, **kw)
    layers.append(layer)
    idx += 2 if layer.norm else 1

  f = filters
  add(name="conv0", kind="conv", k=4, stride=2, cin=3, cout=f, pad="same",
      bias=True, norm=False, act="leaky", kernel_init="normal")
  for i in range(num_downsampling):
    add(name=f"down{i}", kind="conv", k=4, stride=2 if i < 2 else 1, cin=f,
        cout=2 * f, pad="same", bias=False, norm=True, act="leaky",
        kernel_init="normal")
    f *= 2
  add(name="out", kind="conv", k=4, stride=1, cin=f, cout=1, pad="same",
      bias=True, norm=False, act=None, kernel_init="normal")
  return NetSpec(name=name, kind="discriminator", layers=tuple(layers))
