"""Plan-op builders: translate one layer (conv / dgrad / wgrad / norm) into
native kernel launches appended to a ``_C.Plan``.

GEMM weight layout (shared by forward, dgrad (MN-major B) and wgrad output):
``W2[cout_pad][tap * cin + ci]`` with ``tap = r * k + s``; from the TF layouts
    Conv2D kernel [r, s, ci, co]           → permute(3, 0, 1, 2)
    Conv2DTranspose kernel [r, s, co, ci]  → permute(2, 0, 1, 3)
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from ..ops import native as nv

Tap = Tuple[int, int, int, int, int]   # (a-map index, dn, dw, dh, B column)


def pack_conv_weight(kernel_tf: torch.Tensor, transpose_conv: bool = False,
                     rows: Optional[int] = None) -> torch.Tensor:
  """TF kernel → fp32 W2[rows >= cout][k*k*cin] (extra rows are zero)."""
  if transpose_conv:
    w2 = kernel_tf.permute(2, 0, 1, 3)
  else:
    w2 = kernel_tf.permute(3, 0, 1, 2)
  cout = w2.shape[0]
  w2 = w2.reshape(cout, -1).contiguous()
  rows = max(cout, 16) if rows is None else rows
  if rows > cout:
    w2 = torch.cat([w2, w2.new_zeros(rows - cout, w2.shape[1])], 0)
  return w2


def unpack_conv_weight(w2: torch.Tensor, k: int, cin: int, cout: int,
                       transpose_conv: bool = False) -> torch.Tensor:
  w = w2[:cout].reshape(cout, k, k, cin)
  return w.permute(1, 2, 0, 3).contiguous() if transpose_conv else w.permute(1, 2, 3, 0).contiguous()


# --------------------------------------------------------------- tap tables
def taps_conv(k: int, cin: int, origin: int = 0) -> List[Tap]:
  """Stride-1 conv reading input pixel (oh + r + origin, ow + s + origin)."""
  return [(0, 0, s + origin, r + origin, (r * k + s) * cin) for r in range(k) for s in range(k)]


def taps_conv_s2(k: int, cin: int, pad_before: int) -> List[Tap]:
  """Stride-2 conv over 4 parity views: input index 2*o + r - pad_before."""
  taps = []
  for r in range(k):
    for s in range(k):
      ir, ic = r - pad_before, s - pad_before
      a, b = ir % 2, ic % 2
      taps.append((a * 2 + b, 0, (ic - b) // 2, (ir - a) // 2, (r * k + s) * cin))
  return taps


def taps_convT_phase(a: int, b: int, cin: int, k: int = 3, pad_before: int = 0) -> List[Tap]:
  """Transposed conv (stride 2) phase (a, b): out[2m+a, 2n+b] = sum x[m - dy, n - dx] K[ky, kx]
  over taps with (ky - pad_before) ≡ a, (kx - pad_before) ≡ b (mod 2)."""
  taps = []
  for ky in range(k):
    if (ky - pad_before - a) % 2:
      continue
    dy = (ky - pad_before - a) // 2
    for kx in range(k):
      if (kx - pad_before - b) % 2:
        continue
      dx = (kx - pad_before - b) // 2
      taps.append((0, 0, -dx, -dy, (ky * k + kx) * cin))
  return taps


# ------------------------------------------------------------- op builders
def add_conv(plan, name: str, a_maps: Sequence[bytes], taps: Sequence[Tap], k_channels: int,
             w2_bf16: torch.Tensor, out_view, out_h: int, out_w: int, nimg: int, n_channels: int,
             bias: int = 0, act: int = 0, stats: int = 0, b_mn: bool = False,
             block_n: Optional[int] = None, tile: Optional[Tuple[int, int]] = None,
             w_limit: Optional[int] = None,
             phases: Optional[Sequence[Tuple[Sequence[Tap], int]]] = None,
             fp8: bool = False, out_scale: int = 0, halo_map=None):
  """Box-mode implicit GEMM: out[n, h, w, :n_channels] over an out_h x out_w grid.

  k_channels: channels of the A views (K per tap); n_channels: valid output channels."""
  block_n = block_n or nv.pick_block_n(n_channels)
  if b_mn:
    block_n = max(block_n, 64)
  bw, bh = tile or nv.tile_shape(out_h, out_w)
  ptr, sn, sh, sw = out_view
  ph = []
  if phases is not None:   # several tap groups (sub-pixel phases) in ONE launch; out offsets in elements
    taps = []
    for ptaps, off in phases:
      ph.append((len(taps), len(ptaps), int(off)))
      taps.extend(ptaps)
  kblock = 128 if fp8 else 64   # elements per 128-byte K block
  ts_map, ts = b"", []
  if halo_map is not None and tap_share_ok(block_n, fp8, phases, a_maps):
    # halo_map(box_w, box_h) -> tensor map of the same view as a_maps[0] with a halo-sized box
    dws, dhs = [t[2] for t in taps], [t[3] for t in taps]
    box_w, box_h = bw + max(dws) - min(dws), bh + max(dhs) - min(dhs)
    if bw % 8 == 0 and box_w <= 256 and box_h <= 256 and box_w * box_h * 128 <= TS_STAGE_BYTES:
      ts_map = halo_map(box_w, box_h)
      ts = [1, 0, box_w * box_h * 128, box_w, min(dws), min(dhs), box_w * 128 if bw == 8 else 1024]
      if bw != 8 and bh != 1:
        ts_map, ts = b"", []   # 8-row groups must be equally spaced: bw == 8 (pitch) or a single row
  # partial last tile row (bh does not divide out_h): loads are zero-filled, stores masked by h_limit
  return plan.add_conv_gemm(name, list(a_maps), nv.wmat_tmap(w2_bf16, block_n, b_mn),
                            list(taps), k_channels // kblock, out_w // bw, -(-out_h // bh), bw, bh,
                            out_w if w_limit is None else w_limit, ptr, sn, sh, sw, n_channels,
                            bias, act, stats, block_n, b_mn, nimg, ph, out_scale, fp8, ts_map, ts, out_h)


def add_conv_mx(plan, name: str, a_map: bytes, sfa_map: bytes, taps: Sequence[Tap], k_channels: int,
                w8: torch.Tensor, sfb: torch.Tensor, out_view, out_h: int, out_w: int, nimg: int,
                n_channels: int, bias: int = 0, act: int = 0, stats: int = 0,
                tile: Optional[Tuple[int, int]] = None, halo_maps=None):
  """MXFP8 box-mode implicit GEMM (csrc/conv_gemm_mx.cu): E4M3 activations (``a_map``, 128-channel
  boxes) and weights (``w8`` [cout, taps * cin] bytes) with E8M0 scales per 32 K elements —
  ``sfa_map``: 16-byte scale records per pixel; ``sfb``: the weight scales as written by
  ``Plan.add_mx_weight_quant``. 128-channel output tiles. ``halo_maps(box_w, box_h)`` -> (activation,
  scale-record) tensor maps with a halo-sized box enables the tap-shared CTA-pair kernel."""
  _check_mx(name, k_channels, n_channels, len(taps), sfb)
  bw, bh = tile or nv.tile_shape(out_h, out_w)
  ptr, sn, sh, sw = out_view
  ts_map, sfts_map, ts = b"", b"", []
  from ..config import ENGINE
  if halo_maps is not None and ENGINE.tap_share:
    dws, dhs = [t[2] for t in taps], [t[3] for t in taps]
    box_w, box_h = bw + max(dws) - min(dws), bh + max(dhs) - min(dhs)
    if bw == 8 and box_w <= 256 and box_h <= 256 and box_w * box_h * 128 <= TS_STAGE_BYTES:
      ts_map, sfts_map = halo_maps(box_w, box_h)
      ts = [1, 0, box_w * box_h * 128, box_w, min(dws), min(dhs), box_w * 128]
  return plan.add_conv_gemm_mx(name, a_map, sfa_map, nv.wmat_tmap(w8, 128, False), sfb.data_ptr(), list(taps),
                               k_channels // 128, out_w // bw, -(-out_h // bh), bw, bh, out_w, out_h,
                               ptr, sn, sh, sw, n_channels, bias, act, stats, nimg, ts_map, sfts_map, ts)


def add_conv_flat_mx(plan, name: str, a_map: bytes, sfa_map: bytes, taps: Sequence[Tap], k_channels: int,
                     w8: torch.Tensor, sfb: torch.Tensor, out_ptr: int, out_row_stride: int, rows: int,
                     n_channels: int, halo_maps=None):
  """Flat-mode MXFP8 implicit GEMM over ``rows`` consecutive pixels of a zero-ringed domain (dgrad of a
  reflection-padded conv); ``halo_maps(rows_box)`` -> (activation, scale-record) maps for tap sharing."""
  _check_mx(name, k_channels, n_channels, len(taps), sfb)
  tiles = (rows + 127) // 128
  ts_map, sfts_map, ts = b"", b"", []
  from ..config import ENGINE
  if halo_maps is not None and ENGINE.tap_share:
    dws = [t[2] for t in taps]
    span = 128 + max(dws) - min(dws)
    box_rows = -(-((span + 1) // 2) // 8) * 8
    if all(t[3] == 0 for t in taps) and 2 * box_rows * 128 <= TS_STAGE_BYTES and box_rows <= 256:
      ts_map, sfts_map = halo_maps(box_rows)
      ts = [2, box_rows, box_rows * 128, 0, min(dws), 0, 1024]
  return plan.add_conv_gemm_mx(name, a_map, sfa_map, nv.wmat_tmap(w8, 128, False), sfb.data_ptr(), list(taps),
                               k_channels // 128, tiles, 1, 128, 1, rows, 1, out_ptr, 0, 0, out_row_stride,
                               n_channels, 0, 0, 0, 1, ts_map, sfts_map, ts)


def add_mx_weight_quant(plan, name: str, w2_ptr: int, w8_ptr: int, sfb_ptr: int, cout: int, taps: int, cin: int,
                        transposed: bool = False):
  """bf16 conv weights [cout][taps * cin] -> MXFP8 operand + scales: the forward operand (same indexing,
  blocks of 32 input channels) or, ``transposed``, the dgrad operand W^T [cin][taps * cout] (blocks of 32
  output channels)."""
  if transposed:
    return plan.add_mx_weight_quant(name, w2_ptr, w8_ptr, sfb_ptr, cin, taps, cout, 1, cin, taps * cin)
  return plan.add_mx_weight_quant(name, w2_ptr, w8_ptr, sfb_ptr, cout, taps, cin, taps * cin, cin, 1)


def _check_mx(name: str, k_channels: int, n_channels: int, ntaps: int, sfb: torch.Tensor):
  if k_channels % 128 or not 128 <= k_channels <= 512 or n_channels % 128:
    raise ValueError(f"{name}: MXFP8 conv needs cin in 128..512 and cout as multiples of 128")
  if sfb.numel() != n_channels * ntaps * k_channels // 32:
    raise ValueError(f"{name}: weight scale buffer has {sfb.numel()} bytes")


TS_STAGE_BYTES = 34 * 1024   # csrc/conv_gemm_ts.cu: TsCfg::kAStageBytes


def tap_share_ok(block_n: int, fp8: bool, phases, a_maps) -> bool:
  from ..config import ENGINE
  return block_n == 256 and not fp8 and phases is None and len(a_maps) == 1 and ENGINE.tap_share


def add_conv_flat(plan, name: str, a_map: bytes, taps: Sequence[Tap], k_channels: int,
                  w2_bf16: torch.Tensor, out_ptr: int, out_row_stride: int, rows: int,
                  n_channels: int, b_mn: bool = False, block_n: Optional[int] = None,
                  halo_map=None):
  """Flat-mode implicit GEMM over ``rows`` consecutive pixels of a zero-ringed domain
  (taps are constant row offsets): out[row, :n_channels]."""
  block_n = block_n or nv.pick_block_n(n_channels)
  if b_mn:
    block_n = max(block_n, 64)
  tiles = (rows + 127) // 128
  ts_map, ts = b"", []
  if halo_map is not None and tap_share_ok(block_n, False, None, [a_map]):
    # flat rows: the halo of a 128-row tile is [min_dw, 128 + max_dw) rows, loaded as two boxes
    dws = [t[2] for t in taps]
    span = 128 + max(dws) - min(dws)
    box_rows = -(-((span + 1) // 2) // 8) * 8
    if all(t[3] == 0 for t in taps) and 2 * box_rows * 128 <= TS_STAGE_BYTES and box_rows <= 256:
      ts_map = halo_map(box_rows)
      ts = [2, box_rows, box_rows * 128, 0, min(dws), 0, 1024]
  return plan.add_conv_gemm(name, [a_map], nv.wmat_tmap(w2_bf16, block_n, b_mn), list(taps),
                            k_channels // 64, tiles, 1, 128, 1, rows, out_ptr, 0, 0,
                            out_row_stride, n_channels, 0, 0, 0, block_n, b_mn, 1, [], 0, False,
                            ts_map, ts, 1)


def pick_ksplit(base_ctas: int, total_kt: int, slots: int = 148, epilogue_kt: int = 12) -> int:
  """Split-K factor of a wgrad launch, chosen per whole waves of the machine.

  The wgrad kernels need > 114 KB of shared memory, so one CTA is resident per SM and a grid of
  ``base_ctas * ksplit`` CTAs runs in ceil(ctas / 148) waves of ceil(total_kt / ksplit) K tiles
  each plus one (un-overlapped) reduce-add epilogue worth ~12 K tiles (measured: 3x3 256->256,
  12 images: ksplit 8 -> 1 wave, 54 us; ksplit 17 -> 3 waves, 77 us)."""
  best, best_cost = 1, None
  for ks in range(1, max(1, min(total_kt, 4 * slots // max(1, base_ctas) + 1)) + 1):
    waves = -(-base_ctas * ks // slots)
    cost = waves * (-(-total_kt // ks) + epilogue_kt)
    if best_cost is None or cost < best_cost:
      best, best_cost = ks, cost
  return best


def add_wgrad(plan, name: str, dy_map, a_maps: Sequence[bytes], taps: Sequence[Tap],
              kt_w: int, kt_h: int, kbw: int, kbh: int, nimg: int, cin: int, cout: int,
              dw_ptr: int, ldw: int, block_n: Optional[int] = None, target_ctas: Optional[int] = None):
  from ..config import ENGINE
  target_ctas = target_ctas or ENGINE.wgrad_ctas   # 0: wave-aware choice
  block_n = block_n or max(64, nv.pick_block_n(cin))
  m_tiles = (cout + 127) // 128
  n_tiles = (cin + block_n - 1) // block_n
  base = m_tiles * n_tiles * len(taps)
  total_kt = kt_w * kt_h * nimg

  def split(base_ctas):
    if target_ctas:
      return max(1, min(total_kt, (target_ctas + base_ctas - 1) // base_ctas))
    return pick_ksplit(base_ctas, total_kt)

  ksplit = split(base)
  multi = (cin == 64 and 2 <= len(taps) <= 7 and not isinstance(dy_map, (list, tuple))
           and ENGINE.wgrad_multi)
  if multi:   # narrow layers: one CTA accumulates all taps (dY tile loaded once per K tile)
    ksplit = -split(m_tiles)
  # dy_map: one view, or a list of views selected per tap through tap[1] (phase-merged wgrad)
  dy_maps = list(dy_map) if isinstance(dy_map, (list, tuple)) else [dy_map]
  return plan.add_wgrad(name, dy_maps, list(a_maps), list(taps), kt_w, kt_h, kbw, kbh, nimg,
                        cin, cout, dw_ptr, ldw, block_n, ksplit)


def fptr(t: Optional[torch.Tensor]) -> int:
  return 0 if t is None else t.data_ptr()
