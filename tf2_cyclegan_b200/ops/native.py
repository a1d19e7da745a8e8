"""Loader and low-level helpers for the sm_100a extension (``csrc/_C*.so``).

The extension is built in-tree (csrc/build.py, nvcc, ``-gencode
arch=compute_100a,code=sm_100a``) so the ``.so`` travels with the repo
snapshot. On a GPU box a missing extension is a hard error — there is no silent
PyTorch fallback for the native backend.
"""
from __future__ import annotations

import importlib.util
import os
import sys
from typing import Optional, Sequence, Tuple

import torch

_CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "csrc")
_mod = None

ACT = {None: 0, "none": 0, "relu": 1, "leaky": 2, "tanh": 3}
NULL_VIEW = (0, 0, 0, 0)


def so_path() -> str:
  sys.path.insert(0, _CSRC)
  try:
    import build as _b  # csrc/build.py
    return _b.so_path()
  finally:
    sys.path.pop(0)


def build(force: bool = False, verbose: bool = False) -> str:
  sys.path.insert(0, _CSRC)
  try:
    import build as _b
    return _b.build(force=force, verbose=verbose)
  finally:
    sys.path.pop(0)


def load():
  """Import the native module, building it first if the .so is missing."""
  global _mod
  if _mod is not None:
    return _mod
  path = so_path()
  if not os.path.exists(path):
    path = build()
  spec = importlib.util.spec_from_file_location("tf2_cyclegan_b200.csrc._C", path)
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  from ..config import ENGINE
  ENGINE.apply_to_extension(mod)
  _mod = mod
  return mod


def available() -> bool:
  return torch.cuda.is_available() and torch.cuda.get_device_capability()[0] >= 10


# ------------------------------------------------------------------ helpers
def tmap(ptr: int, dims: Sequence[int], strides_elems: Sequence[int], box: Sequence[int],
         elem_bytes: int = 2, dtype: int = 0, swizzle: int = 3) -> bytes:
  """Encode a tiled TMA descriptor. dims/box are innermost-first; strides (in
  elements) are for dims[1:]."""
  return load().encode_tmap(int(ptr), dtype, [int(d) for d in dims],
                            [int(s) * elem_bytes for s in strides_elems],
                            [int(b) for b in box], swizzle)


class Buf:
  """bf16 NHWC activation buffer with an optional halo ring of width ``halo``.

  Zero-initialised once: kernels only ever write interiors (or reflect halos),
  so zero rings needed by dgrad stay zero for the lifetime of the plan."""

  def __init__(self, n: int, h: int, w: int, c: int, halo: int = 0, device="cuda",
               slack_rows: int = 0, storage: Optional[torch.Tensor] = None):
    self.n, self.h, self.w, self.c, self.halo = n, h, w, c, halo
    self.hp, self.wp = h + 2 * halo, w + 2 * halo
    numel = n * self.hp * self.wp * c + slack_rows * self.wp * c
    # ``storage``: the first images of a larger buffer of the same per-image geometry (plans for a
    # smaller batch share the arena of the largest one; halos / zero rings are per image, and the
    # slack row then overlaps the next image's — finite — data, which is all it needs to be)
    assert storage is None or storage.numel() >= numel
    self.storage = torch.zeros(numel, dtype=torch.bfloat16, device=device) if storage is None else storage
    self.t = self.storage[:n * self.hp * self.wp * c].view(n, self.hp, self.wp, c)

  # strides in elements
  @property
  def sn(self): return self.hp * self.wp * self.c
  @property
  def sh(self): return self.wp * self.c
  @property
  def sw(self): return self.c

  def ptr(self, n: int = 0, h: int = 0, w: int = 0, c: int = 0, padded: bool = False) -> int:
    """Address of logical pixel (h, w) of image n (padded=True: physical coords)."""
    off = 0 if padded else self.halo
    return self.t.data_ptr() + 2 * (n * self.sn + (h + off) * self.sh + (w + off) * self.sw + c)

  def view(self, n0: int = 0, padded: bool = False, c0: int = 0):
    return (self.ptr(n0, 0, 0, c0, padded), self.sn, self.sh, self.sw)

  def interior(self, n0: int = 0, n1: Optional[int] = None) -> torch.Tensor:
    p = self.halo
    return self.t[n0:n1, p:p + self.h, p:p + self.w, :]

  # ---- TMA views (box = [64 ch, bw, bh, 1])
  def tmap_padded(self, bw: int, bh: int, n0: int = 0, nimg: Optional[int] = None, cbox: int = 64):
    nimg = self.n - n0 if nimg is None else nimg
    return tmap(self.ptr(n0, padded=True), [self.c, self.wp, self.hp, nimg],
                [self.sw, self.sh, self.sn], [cbox, bw, bh, 1])

  def tmap_interior(self, bw: int, bh: int, n0: int = 0, nimg: Optional[int] = None, cbox: int = 64):
    nimg = self.n - n0 if nimg is None else nimg
    return tmap(self.ptr(n0), [self.c, self.w, self.h, nimg],
                [self.sw, self.sh, self.sn], [cbox, bw, bh, 1])

  def tmap_parity(self, a: int, b: int, bw: int, bh: int, n0: int = 0,
                  nimg: Optional[int] = None):
    """View of pixels (2i + a, 2j + b): element (i, j) = x[2i + a, 2j + b];
    indices outside the image (negative or past the end) are zero-filled by TMA."""
    nimg = self.n - n0 if nimg is None else nimg
    nh = (self.h - a + 1) // 2
    nw = (self.w - b + 1) // 2
    return tmap(self.ptr(n0, a, b), [self.c, nw, nh, nimg],
                [2 * self.sw, 2 * self.sh, self.sn], [64, bw, bh, 1])

  def tmap_flat(self, n0: int = 0, nimg: Optional[int] = None, rows_box: int = 128):
    """All physical pixels of images [n0, n0+nimg) as one [rows, C] matrix."""
    nimg = self.n - n0 if nimg is None else nimg
    rows = nimg * self.hp * self.wp
    return tmap(self.ptr(n0, padded=True), [self.c, rows, 1, 1],
                [self.sw, rows * self.sw, rows * self.sw], [64, rows_box, 1, 1])


class Buf8(Buf):
  """E4M3 (1 byte / element) twin of :class:`Buf`: same logical layout, TMA boxes of 128 channels."""

  def __init__(self, n: int, h: int, w: int, c: int, halo: int = 0, device="cuda", slack_rows: int = 0,
               storage: Optional[torch.Tensor] = None):
    self.n, self.h, self.w, self.c, self.halo = n, h, w, c, halo
    self.hp, self.wp = h + 2 * halo, w + 2 * halo
    numel = n * self.hp * self.wp * c
    self.storage = torch.zeros(numel, dtype=torch.uint8, device=device) if storage is None else storage
    self.t = self.storage[:numel].view(n, self.hp, self.wp, c)

  def ptr(self, n: int = 0, h: int = 0, w: int = 0, c: int = 0, padded: bool = False) -> int:
    off = 0 if padded else self.halo
    return self.t.data_ptr() + (n * self.sn + (h + off) * self.sh + (w + off) * self.sw + c)

  def tmap_padded(self, bw: int, bh: int, n0: int = 0, nimg: Optional[int] = None, cbox: int = 128,
                  swizzle: int = 3):
    """``cbox`` = 128, swizzle 3: E4M3 activations; ``cbox`` = 16, swizzle 0: MX scale records."""
    nimg = self.n - n0 if nimg is None else nimg
    return tmap(self.ptr(n0, padded=True), [self.c, self.wp, self.hp, nimg],
                [self.sw, self.sh, self.sn], [cbox, bw, bh, 1], elem_bytes=1, dtype=2, swizzle=swizzle)


  def tmap_flat(self, n0: int = 0, nimg: Optional[int] = None, rows_box: int = 128, cbox: int = 128,
                swizzle: int = 3):
    nimg = self.n - n0 if nimg is None else nimg
    rows = nimg * self.hp * self.wp
    return tmap(self.ptr(n0, padded=True), [self.c, rows, 1, 1], [self.sw, rows * self.sw, rows * self.sw],
                [cbox, rows_box, 1, 1], elem_bytes=1, dtype=2, swizzle=swizzle)


def tile_shape(h: int, w: int, pixels: int = 128) -> Tuple[int, int]:
  """(bw, bh) with bw * bh == pixels and bw | w: bw is the largest power of two that divides w (capped
  at ``pixels``). bh need NOT divide h: the last tile row of an image is partial — TMA zero-fills the
  rows past the image on load and the epilogue masks them on store (``h_limit``)."""
  bw = 1
  while bw * 2 <= pixels and w % (bw * 2) == 0:
    bw *= 2
  bh = pixels // bw
  if w % bw or bw * bh != pixels or bw > 256 or bh > 256:
    raise ValueError(f"cannot tile {h}x{w} into {pixels}-pixel boxes")
  return bw, bh


def wmat_tmap(w2: torch.Tensor, block_n: int, mn_major: bool) -> bytes:
  """Weights as a [rows, K] bf16 matrix. K-major: box {64, block_n};
  MN-major use (B[k][n] = w2[k, n]): box {64, 64}."""
  rows, k = w2.shape
  if w2.dtype == torch.uint8:   # E4M3 weights: 128 elements per 128-byte K block
    return tmap(w2.data_ptr(), [k, rows], [k], [128, block_n // 2 if block_n >= 128 else block_n],
                elem_bytes=1, dtype=2)
  # wide K-major tiles are fetched as two half-height boxes so that the same descriptor serves
  # the 1-CTA kernel (two loads) and the cta_group::2 kernel (one half per CTA)
  box = [64, 64] if mn_major else [64, block_n // 2 if block_n >= 128 else block_n]
  return tmap(w2.data_ptr(), [k, rows], [k], box)


def pick_block_n(cout: int) -> int:
  if cout <= 16: return 16
  if cout <= 32: return 32
  if cout <= 64: return 64
  if cout <= 128: return 128
  return 256
