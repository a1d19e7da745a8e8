"""Flat parameter storage of the native backend.

Every network owns ONE flat fp32 master buffer (plus Adam m / v, an fp32
gradient buffer and a bf16 compute copy of identical layout). Conv kernels are
stored in the GEMM layout ``W2[rows][taps * cin]`` that forward (K-major B),
dgrad (MN-major B) and wgrad (output) all share, so optimizer, all-reduce and
casts are purely elementwise over the flat buffer — the zero-copy counterpart
of ``NcclAllReduce(num_packs=1)``'s concat.

Small-channel layers (3-channel images) use 8-channel pixels and a "row-window"
packing ``W2[co][r][j(8)][c(8)]`` (one GEMM tap per kernel row, K = 8 pixels x 8
channels); entries that are structural padding are kept at exactly zero by
masking their gradients.

``logical`` ⇄ ``flat`` conversion (TF layouts, Keras object-graph names) is only
needed for initialisation and checkpoints.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from ..models.spec import ConvLayer, NetSpec

ALIGN = 64  # elements; keeps every segment 256-B (fp32) / 128-B (bf16) aligned for TMA


@dataclass
class Segment:
  name: str            # logical variable name (layer_with_weights-i/kernel ...)
  kind: str            # "conv" | "convT" | "rowwin" | "vector"
  offset: int
  shape: Tuple[int, ...]   # flat-layout shape (rows, cols) or (n,)
  logical_shape: Tuple[int, ...]
  layer: Optional[ConvLayer] = None

  @property
  def numel(self) -> int:
    n = 1
    for s in self.shape:
      n *= s
    return n


def is_small_cin(l: ConvLayer) -> bool:
  return l.kind == "conv" and l.cin < 8


def kernel_flat_shape(l: ConvLayer) -> Tuple[str, Tuple[int, int]]:
  if is_small_cin(l):
    return "rowwin", (max(l.cout, 16), l.k * 64)
  rows = max(l.cout, 16)
  return ("convT" if l.kind == "convT" else "conv"), (rows, l.k * l.k * l.cin)


# ------------------------------------------------------------ pack / unpack
def pack_kernel(l: ConvLayer, k_tf: torch.Tensor, fill: float = 0.0) -> torch.Tensor:
  """TF-layout kernel → flat-layout matrix (padding entries = ``fill``)."""
  kind, (rows, cols) = kernel_flat_shape(l)
  out = k_tf.new_full((rows, cols), fill)
  if kind == "rowwin":
    v = out[:l.cout].view(l.cout, l.k, 8, 8)
    v[:, :, :l.k, :l.cin] = k_tf.permute(3, 0, 1, 2)      # [co, r, s, c]
  elif kind == "conv":
    out[:l.cout] = k_tf.permute(3, 0, 1, 2).reshape(l.cout, -1)
  else:  # Conv2DTranspose kernel [r, s, co, ci]
    out[:l.cout] = k_tf.permute(2, 0, 1, 3).reshape(l.cout, -1)
  return out


def unpack_kernel(l: ConvLayer, flat: torch.Tensor) -> torch.Tensor:
  kind, (rows, cols) = kernel_flat_shape(l)
  m = flat.view(rows, cols)
  if kind == "rowwin":
    return m[:l.cout].view(l.cout, l.k, 8, 8)[:, :, :l.k, :l.cin].permute(1, 2, 3, 0).contiguous()
  if kind == "conv":
    return m[:l.cout].view(l.cout, l.k, l.k, l.cin).permute(1, 2, 3, 0).contiguous()
  return m[:l.cout].view(l.cout, l.k, l.k, l.cin).permute(1, 2, 0, 3).contiguous()


class FlatLayout:
  """Segment table of one network."""

  def __init__(self, spec: NetSpec):
    self.spec = spec
    self.segments: List[Segment] = []
    self.by_name: Dict[str, Segment] = {}
    off = 0

    def add(seg: Segment):
      nonlocal off
      self.segments.append(seg)
      self.by_name[seg.name] = seg
      off = seg.offset + seg.numel
      off = (off + ALIGN - 1) // ALIGN * ALIGN

    # kernels first (layer order: the gradient all-reduce buckets are cut at layer boundaries), then
    # every small vector (biases, InstanceNorm gamma / beta) in ONE tail region [vec_lo, numel): on
    # the fused multi-GPU path the kernels are ZeRO-1 sharded (two-shot), the tail is updated by
    # every rank on its own replica (one-shot), so gamma / beta / bias fp32 masters stay current on
    # every rank (csrc/allreduce_adam.cuh)
    for l in spec.layers:
      p = f"layer_with_weights-{l.lw_index}"
      kind, shape = kernel_flat_shape(l)
      add(Segment(f"{p}/kernel", kind, off, shape, l.kernel_shape, l))
    off = (off + 127) // 128 * 128
    self.vec_lo = off
    for l in spec.layers:
      p = f"layer_with_weights-{l.lw_index}"
      if l.bias:
        add(Segment(f"{p}/bias", "vector", off, (l.cout,), (l.cout,), l))
      if l.norm:
        q = f"layer_with_weights-{l.lw_index + 1}"
        add(Segment(f"{q}/gamma", "vector", off, (l.cout,), (l.cout,), l))
        add(Segment(f"{q}/beta", "vector", off, (l.cout,), (l.cout,), l))
    self.numel = (off + 127) // 128 * 128

  def seg(self, l: ConvLayer, what: str) -> Segment:
    idx = l.lw_index + (1 if what in ("gamma", "beta") else 0)
    return self.by_name[f"layer_with_weights-{idx}/{what}"]

  # logical dict (TF layouts) -> flat vector
  def pack(self, logical: Dict[str, torch.Tensor], fill: float = 0.0) -> torch.Tensor:
    flat = torch.full((self.numel,), fill, dtype=torch.float32)
    for s in self.segments:
      t = logical[s.name].detach().float().cpu()
      if s.kind == "vector":
        flat[s.offset:s.offset + s.numel] = t
      else:
        flat[s.offset:s.offset + s.numel] = pack_kernel(s.layer, t, fill).reshape(-1)
    return flat

  def unpack(self, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
    flat = flat.detach().float().cpu()
    out = {}
    for s in self.segments:
      v = flat[s.offset:s.offset + s.numel]
      out[s.name] = v.clone() if s.kind == "vector" else unpack_kernel(s.layer, v)
    return out

  def mask(self) -> torch.Tensor:
    """uint8 [numel]: 1 for real parameters, 0 for structural padding / alignment gaps."""
    ones = {s.name: torch.ones(s.logical_shape) for s in self.segments}
    return (self.pack(ones, 0.0) > 0.5).to(torch.uint8)

  def index_tensor(self, l: ConvLayer) -> torch.Tensor:
    """TF-layout tensor holding the flat index of every kernel element of layer ``l``."""
    s = self.seg(l, "kernel")
    idx = torch.arange(s.offset, s.offset + s.numel, dtype=torch.float64)
    return unpack_kernel(l, idx).to(torch.int64)
