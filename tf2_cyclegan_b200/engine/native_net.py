"""Per-network executor of the native backend: owns the activation / gradient
buffers of one network for a fixed batch and emits its forward / backward kernel
launches into a ``_C.Plan``.

Layouts (all bf16 NHWC):
  * the input of a reflection-padded conv is stored already padded (the
    producing kernel writes the mirrored halo), so those convs are "valid";
  * TF-SAME zero padding is TMA out-of-bounds fill (negative / past-the-end
    coordinates), stride-2 convs read four parity views of the plain tensor;
  * transposed convs are four sub-pixel phase GEMMs writing an interleaved output;
  * dgrad of a reflection-padded conv runs in the zero-ringed "flat" domain
    (taps are constant row offsets) and the halo is folded back by the
    InstanceNorm-backward pass of the layer below;
  * 3-channel tensors use 8-channel pixels; their convs take one GEMM tap per
    kernel row over an overlapping 8-pixel window (16-byte TMA stride).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from ..models.spec import ConvLayer, NetSpec
from ..ops import native as nv
from ..ops import oracle as O
from ..ops.native import ACT, NULL_VIEW, Buf, Buf8
from . import native_ops as no
from .native_params import FlatLayout, is_small_cin

IN_EPS = O.IN_EPS


class NetState:
  """Flat parameter / optimizer / gradient buffers of one network."""

  def __init__(self, spec: NetSpec, logical: Dict[str, torch.Tensor], device,
               grad_buffer: Optional[torch.Tensor] = None,
               wb_buffer: Optional[torch.Tensor] = None):
    self.spec = spec
    self.layout = FlatLayout(spec)
    n = self.layout.numel
    self.n = n
    self.w = self.layout.pack(logical).to(device)
    self.m = torch.zeros(n, device=device)
    self.v = torch.zeros(n, device=device)
    self.g = grad_buffer if grad_buffer is not None else torch.zeros(n, device=device)
    self.wb = wb_buffer if wb_buffer is not None else torch.zeros(n, device=device, dtype=torch.bfloat16)
    self.wb.copy_(self.w.to(torch.bfloat16))
    self.mask = self.layout.mask().to(device)
    self.step = torch.zeros(1, dtype=torch.int32, device=device)
    self.derived: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}   # name -> (bf16 dst, int32 idx)
    self.fp8_layers: List[ConvLayer] = []
    self.fp8_mode = ""
    self.w8: Optional[torch.Tensor] = None

  # ---- FP8 for the residual-block convs: "tensor" = E4M3 with per-tensor power-of-two delayed scaling,
  # "mx" = MXFP8 (E4M3 + one E8M0 scale per 32 K elements, no scaling state)
  def enable_fp8(self, mode: str = "tensor"):
    if self.w8 is not None:
      if mode != self.fp8_mode:
        raise ValueError(f"network already set up for fp8 mode {self.fp8_mode!r}")
      return
    dev = self.w.device
    self.fp8_mode = mode
    self.fp8_layers = [l for l in self.spec.layers if fp8_ok(l)]
    L = max(1, len(self.fp8_layers))
    self.w8 = torch.zeros(self.n, dtype=torch.uint8, device=dev)
    # MX weight scales: one byte per (row, 32 K elements), per layer — of the forward operand [cout][taps * cin]
    # and of the dgrad operand W^T [cin][taps * cout] (its own E4M3 copy: the blocks run over output channels)
    nsf = [self.seg(l, "kernel").numel // 32 if mode == "mx" else 0 for l in self.fp8_layers]
    self.w_sf = [torch.zeros(k, dtype=torch.uint8, device=dev) for k in nsf]
    self.w_sft = [torch.zeros(k, dtype=torch.uint8, device=dev) for k in nsf]
    self.w8t = torch.zeros(self.n if mode == "mx" else 0, dtype=torch.uint8, device=dev)
    self.act_amax = torch.zeros(L, device=dev)
    self.w_amax = torch.zeros(L, device=dev)
    self.act_scale = torch.ones(L, device=dev)
    self.w_scale = torch.ones(L, device=dev)
    self.out_scale = torch.ones(L, device=dev)

  def fp8_slot(self, l: ConvLayer) -> int:
    return self.fp8_layers.index(l)

  def kernel_fp8(self, l: ConvLayer) -> torch.Tensor:
    s = self.seg(l, "kernel")
    return self.w8[s.offset:s.offset + s.numel].view(*s.shape)

  def kernel_fp8_t(self, l: ConvLayer) -> torch.Tensor:
    """MXFP8 dgrad operand W^T as [cin, taps * cout] E4M3 bytes."""
    s = self.seg(l, "kernel")
    return self.w8t[s.offset:s.offset + s.numel].view(l.cin, l.k * l.k * l.cout)

  def emit_fp8_refresh(self, plan, tag: str):
    """Start-of-step bookkeeping: weight amax -> scales (activations: previous step's amax) ->
    E4M3 weight copy. Reads the bf16 compute copy, which is complete on every rank."""
    if not self.fp8_layers:
      return
    if self.fp8_mode == "mx":
      for i, l in enumerate(self.fp8_layers):
        seg = self.seg(l, "kernel")
        taps, src = l.k * l.k, self.wb.data_ptr() + 2 * seg.offset
        no.add_mx_weight_quant(plan, f"fp8.{tag}.{l.name}.mx_w_quant", src, self.w8.data_ptr() + seg.offset,
                               self.w_sf[i].data_ptr(), l.cout, taps, l.cin)
        no.add_mx_weight_quant(plan, f"fp8.{tag}.{l.name}.mx_wt_quant", src, self.w8t.data_ptr() + seg.offset,
                               self.w_sft[i].data_ptr(), l.cout, taps, l.cin, transposed=True)
      return
    offs = [self.seg(l, "kernel").offset for l in self.fp8_layers]
    nums = [self.seg(l, "kernel").numel for l in self.fp8_layers]
    args = (offs, nums, self.wb.data_ptr(), self.w8.data_ptr(), self.act_amax.data_ptr(),
            self.w_amax.data_ptr(), self.act_scale.data_ptr(), self.w_scale.data_ptr(),
            self.out_scale.data_ptr(), 16.0)
    plan.add_memset(f"zero.{tag}.fp8_w_amax", self.w_amax.data_ptr(), self.w_amax.numel() * 4)
    plan.add_fp8_op(f"fp8.{tag}.w_amax", 0, *args)
    plan.add_fp8_op(f"fp8.{tag}.scales", 1, *args)
    plan.add_fp8_op(f"fp8.{tag}.w_quant", 2, *args)

  def seg(self, l: ConvLayer, what: str):
    return self.layout.seg(l, what)

  def kernel_bf16(self, l: ConvLayer) -> torch.Tensor:
    s = self.seg(l, "kernel")
    return self.wb[s.offset:s.offset + s.numel].view(*s.shape)

  def wptr(self, l: ConvLayer, what: str) -> int:
    return self.w.data_ptr() + 4 * self.seg(l, what).offset

  def gptr(self, l: ConvLayer, what: str) -> int:
    return self.g.data_ptr() + 4 * self.seg(l, what).offset

  def add_derived(self, name: str, idx: torch.Tensor) -> torch.Tensor:
    """bf16 matrix gathered from the master weights through ``idx`` (-1 = zero)."""
    if name in self.derived:      # shared by every plan (batch size) of this network
      return self.derived[name][0]
    idx32 = idx.to(torch.int32).contiguous().to(self.w.device)
    dst = torch.zeros(idx.shape, dtype=torch.bfloat16, device=self.w.device)
    self.derived[name] = (dst, idx32)
    return dst

  def refresh_derived_ops(self, plan, tag: str):
    # source = the bf16 compute copy: it is complete on every rank, whereas the fp32 master is
    # only current inside the rank's own optimizer shard on the fused (ZeRO-1) path
    for name, (dst, idx) in self.derived.items():
      plan.add_gather_bf16(f"{tag}.gather.{name}", dst.data_ptr(), self.wb.data_ptr(), idx.data_ptr(),
                           idx.numel(), True)

  def refresh_derived_now(self):
    for dst, idx in self.derived.values():
      src = torch.cat([self.wb, self.wb.new_zeros(1)])
      dst.copy_(src[idx.long().clamp_min(-1)].view_as(dst))


def fp8_ok(l: ConvLayer) -> bool:
  """Layers that run their forward conv in FP8: the residual-block 3x3 convs."""
  return l.kind == "conv" and l.res in ("a", "b") and l.cin % 128 == 0 and l.cout == 256


def _out_hw(l: ConvLayer, h: int, w: int) -> Tuple[int, int]:
  return l.out_hw(h, w)


class NetExec:
  """Buffers + op emission for ``nb`` images of one network at resolution ``size`` (int or (h, w))."""

  def __init__(self, spec: NetSpec, state: NetState, nb: int, size, device, tag: str,
               fp8="", pool: Optional[dict] = None):
    """``fp8``: "" / False = bf16, "tensor" (or True) = per-tensor E4M3, "mx" = MXFP8 residual forward convs."""
    self.spec, self.st, self.nb, self.size, self.tag = spec, state, nb, size, tag
    fp8 = "tensor" if fp8 is True else (fp8 or "")
    self.fp8 = fp8 if any(fp8_ok(l) for l in spec.layers) else ""
    self.mx = self.fp8 == "mx"
    if self.fp8:
      state.enable_fp8(self.fp8)
    self.layers = list(spec.layers)
    L = len(self.layers)
    dev = device
    # activation / gradient arenas are shared between the plans of different batch sizes: the
    # partial last batch of an epoch (or a b = 1 plot pass) reuses the first images of the arenas
    # of the full batch instead of doubling the memory (pool: role -> largest buffer so far)
    self.pool = pool
    self.device = dev
    # resolution of every layer input / output
    self.in_hw: List[Tuple[int, int]] = []
    h, w = (size, size) if isinstance(size, int) else size
    for l in self.layers:
      self.in_hw.append((h, w))
      h, w = _out_hw(l, h, w)
    self.out_hw = [(_out_hw(l, *hw)) for l, hw in zip(self.layers, self.in_hw)]

    def cpad(c):  # channel storage (3 -> 8)
      return 8 if c < 8 else c

    # ---- forward activations: inp[j] feeds layer j; inp[L] is the network output
    self.inp: List[Buf] = []
    for j, l in enumerate(self.layers):
      hh, ww = self.in_hw[j]
      halo = l.k // 2 if l.pad == "reflect" else (1 if is_small_cin(l) and l.pad == "same" else 0)
      slack = 1 if is_small_cin(l) else 0
      self.inp.append(self._buf(f"inp{j}", nb, hh, ww, cpad(l.cin), halo=halo, slack_rows=slack))
    ho, wo = self.out_hw[-1]
    self.inp.append(self._buf("out", nb, ho, wo, cpad(self.layers[-1].cout)))
    # E4M3 twins of the inputs of the FP8 layers (written by the producing IN-apply kernel)
    self.inp8: Dict[int, Buf8] = {}
    self.sf8: Dict[int, Buf8] = {}
    if self.fp8:
      for j, l in enumerate(self.layers):
        if fp8_ok(l):
          hh, ww = self.in_hw[j]
          self.inp8[j] = self._buf(f"inp8_{j}", nb, hh, ww, l.cin, halo=l.k // 2, cls=Buf8)
          if self.mx:   # 16-byte scale records per pixel (byte i = channels [32 i, 32 i + 32))
            self.sf8[j] = self._buf(f"sf8_{j}", nb, hh, ww, 16, halo=l.k // 2, cls=Buf8)
    self.raw: List[Optional[Buf]] = []
    self.stats: List[Optional[torch.Tensor]] = []
    self.bstats: List[Optional[torch.Tensor]] = []
    stat_sizes = [nb * l.cout * 2 if l.norm else 0 for l in self.layers]
    self.stat_arena = torch.zeros(sum(stat_sizes) + 8, device=dev)     # zeroed every step
    self.bstat_arena = torch.zeros(sum(stat_sizes) + 8, device=dev)    # zeroed before every backward pass
    off = 0
    for j, l in enumerate(self.layers):
      ho, wo = self.out_hw[j]
      if l.norm:
        self.raw.append(self._buf(f"raw{j}", nb, ho, wo, l.cout))
        self.stats.append(self.stat_arena[off:off + stat_sizes[j]].view(nb, l.cout, 2))
        self.bstats.append(self.bstat_arena[off:off + stat_sizes[j]].view(nb, l.cout, 2))
        off += stat_sizes[j]
      else:
        self.raw.append(None)
        self.stats.append(None)
        self.bstats.append(None)

    # ---- backward buffers (allocated lazily by enable_backward)
    self.dY: List[Optional[Buf]] = [None] * L      # grad wrt conv output (zero ring for flat dgrad)
    self.dY8: Dict[int, Buf8] = {}                 # MXFP8 twins of dY (residual-block convs, --dtype mxfp8)
    self.dYsf: Dict[int, Buf8] = {}
    self.dyt: List[Optional[Buf]] = [None] * L     # g * act' scratch (doubles as the residual-sum gradient)
    self.gin: List[Optional[Buf]] = [None] * L     # grad wrt layer input, as produced by the layer's dgrad
    self.gout_final: Optional[Buf] = None          # grad wrt the network output (images / logits)
    self._bwd = False
    self._special_weights()

  def _buf(self, role: str, n: int, h: int, w: int, c: int, halo: int = 0, slack_rows: int = 0, cls=Buf):
    if self.pool is None:
      return cls(n, h, w, c, halo=halo, device=self.device, slack_rows=slack_rows)
    key = (self.tag, role, h, w, c, halo, slack_rows, cls.__name__)
    big = self.pool.get(key)
    if big is not None and big.n >= n:
      return cls(n, h, w, c, halo=halo, device=self.device, slack_rows=slack_rows, storage=big.storage)
    b = cls(n, h, w, c, halo=halo, device=self.device, slack_rows=slack_rows)
    self.pool[key] = b
    return b

  # --------------------------------------------------------------- weights
  def _special_weights(self):
    """Derived weight matrices for the small-channel layers' dgrad."""
    st = self.st
    self.w_dgrad: Dict[int, torch.Tensor] = {}
    self.w_rowgemm: Dict[int, torch.Tensor] = {}
    self.q_buf: Dict[int, torch.Tensor] = {}
    for j, l in enumerate(self.layers):
      if l.kind == "conv" and l.cout < 8 and l.cout > 1:
        # generator final conv (Cout = 3): dgrad B (K-major) Wfd[ci][r*64 + jj*8 + co] = K[r][k-1-jj][ci][co]
        idx_tf = st.layout.index_tensor(l)                       # [k, k, cin, cout]
        d = torch.full((l.cin, l.k, 8, 8), -1, dtype=torch.int64)
        for jj in range(l.k):
          d[:, :, jj, :l.cout] = idx_tf[:, l.k - 1 - jj, :, :].permute(1, 0, 2)
        self.w_dgrad[j] = st.add_derived(f"{self.tag}.L{j}.dgradW", d.reshape(l.cin, l.k * 64))
        # forward as a row-GEMM: Q[q][s*4 + co] = sum_{r,ci} a_pad[q + r*P][ci] K[r][s][ci][co]
        q = torch.full((32, l.k, l.cin), -1, dtype=torch.int64)
        for ss in range(l.k):
          q[ss * 4:ss * 4 + l.cout] = idx_tf[:, ss, :, :].permute(2, 0, 1)        # [co, r, ci]
        self.w_rowgemm[j] = st.add_derived(f"{self.tag}.L{j}.rowW", q.reshape(32, l.k * l.cin))
      elif is_small_cin(l) and l.stride == 2:
        # PatchGAN conv0 (Cin = 3): dgrad B (K-major) W[c(16)][(r*k+s)*cout + co] = K[r][s][c][co]
        idx_tf = st.layout.index_tensor(l)                       # [k, k, cin, cout]
        d = torch.full((16, l.k * l.k * l.cout), -1, dtype=torch.int64)
        d[:l.cin] = idx_tf.permute(2, 0, 1, 3).reshape(l.cin, -1)
        self.w_dgrad[j] = st.add_derived(f"{self.tag}.L{j}.dgradW", d)
    st.refresh_derived_now()

  def enable_backward(self):
    if self._bwd:
      return
    self._bwd = True
    dev = self.inp[0].t.device
    nb = self.nb
    L = len(self.layers)
    for j, l in enumerate(self.layers):
      ho, wo = self.out_hw[j]
      hi, wi = self.in_hw[j]
      c_out = 8 if l.cout < 8 else l.cout
      flat_dgrad = l.kind == "conv" and l.pad == "reflect" and not is_small_cin(l)
      halo = l.k // 2 if flat_dgrad else 0
      self.dY[j] = self._buf(f"dY{j}", nb, ho, wo, c_out, halo=halo, slack_rows=1 if c_out == 8 else 0)
      if l.norm:
        self.dyt[j] = self._buf(f"dyt{j}", nb, ho, wo, l.cout)
      if self.mx and fp8_ok(l) and flat_dgrad and l.norm:
        # MXFP8 twin of dY for the dgrad (written by the IN-backward-apply kernel): bytes + scale records
        self.dY8[j] = self._buf(f"dY8_{j}", nb, ho, wo, l.cout, halo=halo, cls=Buf8)
        self.dYsf[j] = self._buf(f"dYsf_{j}", nb, ho, wo, 16, halo=halo, cls=Buf8)
      if j > 0 or True:
        c_in = 8 if l.cin < 8 else l.cin
        self.gin[j] = self._buf(f"gin{j}", nb, hi, wi, c_in, halo=halo) if not is_small_cin(l) else None
    ho, wo = self.out_hw[-1]
    self.gout_final = self._buf("gout", nb, ho, wo, self.inp[-1].c)
    # scratch for the swapped wgrad of small-Cout reflect convs (generator final layer)
    self.wgrad_scratch: Dict[int, torch.Tensor] = {}
    for j, l in enumerate(self.layers):
      if j in self.w_dgrad and not is_small_cin(l):
        self.wgrad_scratch[j] = torch.zeros(128, l.k * 64, device=dev)

  # --------------------------------------------------------------- helpers
  def _small_cin_maps(self, j: int, n0: int, nimg: int, bw: int, bh: int):
    """Overlapping-window A views + taps of a 3-channel-input conv (8-channel pixels)."""
    l = self.layers[j]
    src = self.inp[j]
    ho, wo = self.out_hw[j]
    if l.stride == 1:
      maps = [nv.tmap(src.ptr(n0, padded=True), [64, wo, src.hp, nimg], [8, src.sh, src.sn],
                      [64, bw, bh, 1])]
      taps = [(0, 0, 0, r, r * 64) for r in range(l.k)]
    else:
      maps = []
      for a in (0, 1):
        nrows = (src.hp - a + 1) // 2
        maps.append(nv.tmap(src.ptr(n0, padded=True) + 2 * a * src.sh, [64, wo, nrows, nimg],
                            [16, 2 * src.sh, src.sn], [64, bw, bh, 1]))
      taps = [(r % 2, 0, 0, r // 2, r * 64) for r in range(l.k)]
    return maps, taps

  def _pad_before(self, j: int) -> int:
    l = self.layers[j]
    hi, _ = self.in_hw[j]
    if l.kind == "convT":
      return 0
    return O.same_pad_amounts(hi, l.k, l.stride)[0]

  # --------------------------------------------------------------- forward
  def forward(self, plan, n0: int, nimg: int, first: int = 0):
    """Emit the forward pass for images [n0, n0 + nimg)."""
    st = self.st
    L = len(self.layers)
    for j in range(first, L):
      l = self.layers[j]
      src = self.inp[j]
      hi, wi = self.in_hw[j]
      ho, wo = self.out_hw[j]
      dst = self.raw[j] if l.norm else self.inp[j + 1]
      name = f"{self.tag}.{l.name}.fwd[{n0}:{n0 + nimg}]"
      bias = st.wptr(l, "bias") if l.bias else 0
      act = 0 if l.norm else ACT[l.act]
      stats = self.stats[j][n0:].data_ptr() if l.norm else 0
      w2 = st.kernel_bf16(l)
      if l.kind == "convT":
        bw, bh = nv.tile_shape(hi, wi)
        amap = src.tmap_interior(bw, bh, n0, nimg)
        view = (dst.ptr(n0), dst.sn, 2 * dst.sh, 2 * dst.sw)
        phases = [(no.taps_convT_phase(a, b, l.cin), a * dst.sh + b * dst.sw) for a in (0, 1) for b in (0, 1)]
        no.add_conv(plan, name, [amap], [], l.cin, w2, view, hi, wi, nimg, l.cout, bias, act, stats,
                    phases=phases)
      elif j in self.w_rowgemm:
        # small-Cout reflect conv (generator output layer): one GEMM tap per kernel ROW over the flat
        # padded domain (A traffic / k), then a k-tap horizontal gather fused with bias + tanh
        if j not in self.q_buf:
          need = (self.nb * src.hp * src.wp + 256) * 32
          qkey = (self.tag, f"q{j}", src.hp, src.wp)
          shared = self.pool.get(qkey) if self.pool is not None else None
          if shared is None or shared.numel() < need:
            shared = torch.zeros(need, dtype=torch.bfloat16, device=src.t.device)
            if self.pool is not None:
              self.pool[qkey] = shared
          self.q_buf[j] = shared
        qb = self.q_buf[j]
        rows = nimg * src.hp * src.wp
        qptr = qb.data_ptr() + 2 * 32 * n0 * src.hp * src.wp
        taps = [(0, 0, r * src.wp, 0, r * l.cin) for r in range(l.k)]
        no.add_conv_flat(plan, name, src.tmap_flat(n0, nimg), taps, l.cin, self.w_rowgemm[j], qptr, 32,
                         rows, 32, block_n=32)
        plan.add_tap_gather(f"{self.tag}.{l.name}.gather[{n0}:{n0 + nimg}]", qptr, bias, dst.view(n0), act,
                            l.k, src.hp, src.wp, nimg, ho, wo)
      else:
        bw, bh = nv.tile_shape(ho, wo)
        if is_small_cin(l):
          maps, taps = self._small_cin_maps(j, n0, nimg, bw, bh)
          kch = 64
        elif j in self.sf8:
          # MXFP8 forward: block-scaled E4M3 activations x weights (scales applied inside the MMA)
          slot = st.fp8_slot(l)
          a8, s8 = self.inp8[j], self.sf8[j]
          tile, halo = None, None
          if ho % 16 == 0 and wo % 8 == 0:   # 8 x 16 tiles: one halo tile per channel block (CTA-pair kernel)
            tile, (bw, bh) = (8, 16), (8, 16)
            halo = lambda w_, h_: (a8.tmap_padded(w_, h_, n0, nimg),
                                   s8.tmap_padded(w_, h_, n0, nimg, cbox=16, swizzle=0))
          no.add_conv_mx(plan, name + ".mxfp8", a8.tmap_padded(bw, bh, n0, nimg),
                         s8.tmap_padded(bw, bh, n0, nimg, cbox=16, swizzle=0),
                         no.taps_conv(l.k, l.cin), l.cin, st.kernel_fp8(l), st.w_sf[slot], dst.view(n0), ho, wo,
                         nimg, l.cout, bias, act, stats, tile=tile, halo_maps=halo)
          maps = None
        elif j in self.inp8:
          # FP8 forward: E4M3 activations x E4M3 weights, fp32 accumulate, de-scaled in the epilogue
          slot = st.fp8_slot(l)
          no.add_conv(plan, name + ".fp8", [self.inp8[j].tmap_padded(bw, bh, n0, nimg)],
                      no.taps_conv(l.k, l.cin), l.cin, st.kernel_fp8(l), dst.view(n0), ho, wo, nimg, l.cout,
                      bias, act, stats, fp8=True, out_scale=st.out_scale[slot:].data_ptr())
          maps = None
        elif l.pad == "reflect":
          maps, taps, kch = [src.tmap_padded(bw, bh, n0, nimg)], no.taps_conv(l.k, l.cin), l.cin
        elif l.stride == 2:
          maps = [src.tmap_parity(a, b, bw, bh, n0, nimg) for a in (0, 1) for b in (0, 1)]
          taps, kch = no.taps_conv_s2(l.k, l.cin, self._pad_before(j)), l.cin
        else:
          maps = [src.tmap_interior(bw, bh, n0, nimg)]
          taps, kch = no.taps_conv(l.k, l.cin, origin=-self._pad_before(j)), l.cin
        if maps is not None:
          tile, halo = None, None
          if (len(maps) == 1 and l.stride == 1 and not is_small_cin(l) and l.kind == "conv"
              and nv.pick_block_n(l.cout) == 256 and ho % 16 == 0 and wo % 8 == 0):
            # wide stride-1 convs: 8 x 16 pixel tiles whose taps share one halo tile (conv_gemm_ts.cu)
            tile = (8, 16)
            view_of = src.tmap_padded if l.pad == "reflect" else src.tmap_interior
            halo = lambda w_, h_, f=view_of: f(w_, h_, n0, nimg)
            maps = [view_of(8, 16, n0, nimg)]
          no.add_conv(plan, name, maps, taps, kch, w2, dst.view(n0), ho, wo, nimg, l.cout, bias, act, stats,
                      tile=tile, halo_map=halo)
      if l.norm:
        nxt = self.inp[j + 1]
        skip = self.inp[j - 1].view(n0) if l.res == "b" else NULL_VIEW
        dst8, scale8, amax8, sfv = NULL_VIEW, 0, 0, NULL_VIEW
        if (j + 1) in self.sf8:
          dst8, sfv = self.inp8[j + 1].view(n0), self.sf8[j + 1].view(n0)
        elif (j + 1) in self.inp8:
          b8 = self.inp8[j + 1]
          slot = st.fp8_slot(self.layers[j + 1])
          dst8 = b8.view(n0)
          scale8 = st.act_scale[slot:].data_ptr()
          amax8 = st.act_amax[slot:].data_ptr()
        plan.add_in_apply(f"{self.tag}.{l.name}.in_apply[{n0}:{n0 + nimg}]", self.raw[j].view(n0),
                          stats, st.wptr(l, "gamma"), st.wptr(l, "beta"), 1.0 / (ho * wo), IN_EPS,
                          ACT[l.act], skip, nxt.view(n0), nxt.halo, nimg, ho, wo, l.cout,
                          dst8, scale8, amax8, sfv)

  # -------------------------------------------------------------- backward
  def backward(self, plan, n0: int, nimg: int, wgrad: bool, stop_layer: int = 0,
               input_grad_target: Optional[Tuple[Buf, int]] = None,
               gout_view=None, gout_is_dy: bool = False,
               lanes: Optional[Tuple[int, int]] = None, join_wgrad: bool = True,
               on_layer_done=None):
    """Emit the backward pass for images [n0, n0 + nimg).

    ``gout_view``: view of the gradient wrt the network output (default: self.gout_final).
    ``gout_is_dy``: the output layer is linear (PatchGAN logits) so gout is already dY.
    ``wgrad``: accumulate parameter gradients into the flat fp32 gradient buffer.
    ``input_grad_target``: (Buf, image offset) receiving the gradient wrt the input image
    (only for the adversarial pass through a discriminator); otherwise the first layer
    has no dgrad. Layers below ``stop_layer`` are skipped entirely.
    ``lanes``: (main lane, wgrad lane): weight-gradient GEMMs are issued on their own stream —
    nothing downstream of a layer depends on its wgrad — so they fill the tail waves of the
    dgrad / normalisation chain; the caller joins the wgrad lane before the optimizer.
    ``on_layer_done(j)``: called after every op that writes layer j's parameter gradients has
    been emitted (layers are visited last to first): hook for bucketed gradient all-reduce."""
    main_lane, wg_lane = lanes if lanes is not None else (None, None)
    self.enable_backward()
    st = self.st
    L = len(self.layers)
    g1 = (gout_view if gout_view is not None else self.gout_final.view(n0), 0)
    g2 = NULL_VIEW
    pending_skip = None
    prev_j = None
    for j in range(L - 1, stop_layer - 1, -1):
      if prev_j is not None and on_layer_done is not None:
        on_layer_done(prev_j)
      prev_j = j
      l = self.layers[j]
      hi, wi = self.in_hw[j]
      ho, wo = self.out_hw[j]
      tagj = f"{self.tag}.{l.name}"
      rng = f"[{n0}:{n0 + nimg}]"
      dY = self.dY[j]
      last = (j == L - 1)
      # ---- 1. gradient wrt the conv output
      if l.norm:
        dyt = self.dyt[j]
        plan.add_in_bwd_reduce(f"{tagj}.in_bwd_reduce{rng}", g1[0], g1[1], g2, self.raw[j].view(n0),
                               self.stats[j][n0:].data_ptr(), st.wptr(l, "gamma"), st.wptr(l, "beta"),
                               1.0 / (ho * wo), IN_EPS, ACT[l.act], dyt.view(n0),
                               self.bstats[j][n0:].data_ptr(), nimg, ho, wo, l.cout)
        plan.add_in_bwd_apply(f"{tagj}.in_bwd_apply{rng}", dyt.view(n0), self.raw[j].view(n0),
                              self.stats[j][n0:].data_ptr(), self.bstats[j][n0:].data_ptr(),
                              st.wptr(l, "gamma"), 1.0 / (ho * wo), IN_EPS, dY.view(n0),
                              st.gptr(l, "gamma") if wgrad else 0, st.gptr(l, "beta") if wgrad else 0,
                              nimg, ho, wo, l.cout,
                              self.dY8[j].view(n0) if j in self.dY8 else NULL_VIEW,
                              self.dYsf[j].view(n0) if j in self.dY8 else NULL_VIEW)
        if l.res == "b":
          pending_skip = dyt.view(n0)   # == gradient wrt the block output (IN-b has no activation)
      elif last and gout_is_dy:
        dY_view_override = g1[0]
      else:
        plan.add_act_bwd(f"{tagj}.act_bwd{rng}", g1[0], g1[1], g2, self.inp[j + 1].view(n0),
                         ACT[l.act], dY.view(n0), st.gptr(l, "bias") if (wgrad and l.bias) else 0,
                         l.cout, nimg, ho, wo, dY.c)
      g2 = NULL_VIEW
      need_dgrad = j > stop_layer or (j == 0 and input_grad_target is not None)
      if j == 0 and input_grad_target is None:
        need_dgrad = False

      # ---- PatchGAN logits layer (Cout = 1): GEMV-shaped SIMT kernels
      if l.cout == 1:
        gd = dY_view_override if (last and gout_is_dy) else dY.view(n0)
        gin = self.gin[j]
        w_row0 = st.kernel_bf16(l).data_ptr()
        plan.add_d4_bwd(f"{tagj}.bwd{rng}", gd, self.inp[j].view(n0), w_row0,
                        gin.view(n0) if need_dgrad else NULL_VIEW,
                        st.gptr(l, "kernel") if wgrad else 0, st.gptr(l, "bias") if (wgrad and l.bias) else 0,
                        self._pad_before(j), nimg, hi, wi, l.cin, bool(need_dgrad), bool(wgrad))
        g1 = (gin.view(n0), 0)
        continue

      # ---- 2. wgrad
      if wgrad:
        if wg_lane is not None:
          ev = plan.add_event_record(f"sync.rec.{tagj}.dY", main_lane)
          plan.set_lane(wg_lane)
          plan.add_event_wait(f"sync.wait.{tagj}.dY", wg_lane, ev)
          self._wgrad(plan, j, n0, nimg, f"{tagj}.wgrad{rng}")
          plan.set_lane(main_lane)
        else:
          self._wgrad(plan, j, n0, nimg, f"{tagj}.wgrad{rng}")
      # ---- 3. dgrad
      if need_dgrad:
        g1 = self._dgrad(plan, j, n0, nimg, f"{tagj}.dgrad{rng}", input_grad_target if j == 0 else None)
        if l.res == "a" and pending_skip is not None:
          g2 = pending_skip
          pending_skip = None
    if prev_j is not None and on_layer_done is not None:
      on_layer_done(prev_j)
    if wgrad and wg_lane is not None and join_wgrad:
      ev = plan.add_event_record(f"sync.rec.{self.tag}.wgrad_done", wg_lane)
      plan.add_event_wait(f"sync.wait.{self.tag}.wgrad_done", main_lane, ev)

  def _wgrad(self, plan, j: int, n0: int, nimg: int, name: str):
    l = self.layers[j]
    st = self.st
    hi, wi = self.in_hw[j]
    ho, wo = self.out_hw[j]
    dY, src = self.dY[j], self.inp[j]
    seg = st.seg(l, "kernel")
    dw, ldw = st.gptr(l, "kernel"), seg.shape[1]
    if l.kind == "convT":
      kbw, kbh = nv.tile_shape(hi, wi, 64)
      amap = src.tmap_interior(kbw, kbh, n0, nimg)
      dy_maps, taps = [], []
      for a in (0, 1):
        for b in (0, 1):
          dy_maps.append(dY.tmap_parity(a, b, kbw, kbh, n0, nimg))
          taps += [(t[0], len(dy_maps) - 1, t[2], t[3], t[4]) for t in no.taps_convT_phase(a, b, l.cin)]
      no.add_wgrad(plan, name, dy_maps, [amap], taps, wi // kbw, -(-hi // kbh), kbw, kbh, nimg, l.cin,
                   l.cout, dw, ldw)
      return
    if j in self.w_dgrad and not is_small_cin(l):
      # small-Cout reflect conv (generator final layer): swapped operands in the flat domain.
      #   dWT[ci][r*64 + jj*8 + co] = sum_u a_pad[u][ci] * dYz[u + (p - r) * P - p + jj][co]
      P = src.wp
      p = l.k // 2
      rows = nimg * src.hp * src.wp
      m_map = src.tmap_flat(n0, nimg, rows_box=64)
      b_map = nv.tmap(dY.ptr(n0, padded=True), [64, rows, 1, 1], [8, rows * 8, rows * 8], [64, 64, 1, 1])
      taps = [(0, 0, (p - r) * P - p, 0, r * 64) for r in range(l.k)]
      scratch = self.wgrad_scratch[j]
      plan.add_memset(f"{name}.zero", scratch.data_ptr(), scratch.numel() * 4)
      no.add_wgrad(plan, name, m_map, [b_map], taps, (rows + 63) // 64, 1, 64, 1, 1, 64, l.cin,
                   scratch.data_ptr(), scratch.shape[1])
      _, idx = st.derived[f"{self.tag}.L{j}.dgradW"]
      plan.add_scatter_add(f"{name}.scatter", st.g.data_ptr(), scratch.data_ptr(), idx.data_ptr(), idx.numel())
      return
    kbw, kbh = nv.tile_shape(ho, wo, 64)
    dy_map = dY.tmap_interior(kbw, kbh, n0, nimg)
    if is_small_cin(l):
      maps, taps = self._small_cin_maps(j, n0, nimg, kbw, kbh)
      cin = 64
    elif l.pad == "reflect":
      maps, taps, cin = [src.tmap_padded(kbw, kbh, n0, nimg)], no.taps_conv(l.k, l.cin), l.cin
    elif l.stride == 2:
      maps = [src.tmap_parity(a, b, kbw, kbh, n0, nimg) for a in (0, 1) for b in (0, 1)]
      taps, cin = no.taps_conv_s2(l.k, l.cin, self._pad_before(j)), l.cin
    else:
      maps = [src.tmap_interior(kbw, kbh, n0, nimg)]
      taps, cin = no.taps_conv(l.k, l.cin, origin=-self._pad_before(j)), l.cin
    no.add_wgrad(plan, name, dy_map, maps, taps, wo // kbw, -(-ho // kbh), kbw, kbh, nimg, cin, l.cout,
                 dw, ldw)

  def _dgrad(self, plan, j: int, n0: int, nimg: int, name: str, input_target):
    """Returns (view, fold) of the gradient wrt the layer input."""
    l = self.layers[j]
    st = self.st
    hi, wi = self.in_hw[j]
    ho, wo = self.out_hw[j]
    dY = self.dY[j]
    w2 = st.kernel_bf16(l)
    if l.kind == "convT":
      gin = self.gin[j]
      bw, bh = nv.tile_shape(hi, wi)
      maps = [dY.tmap_parity(a, b, bw, bh, n0, nimg) for a in (0, 1) for b in (0, 1)]
      no.add_conv(plan, name, maps, no.taps_conv_s2(l.k, l.cin, 0), l.cout, w2, gin.view(n0), hi, wi,
                  nimg, l.cin, b_mn=True)
      return (gin.view(n0), 0)
    if l.pad == "reflect":
      gin = self.gin[j]
      p = l.k // 2
      P = dY.wp
      rows = nimg * dY.hp * dY.wp
      if l.cout < 8:
        amap = nv.tmap(dY.ptr(n0, padded=True), [64, rows, 1, 1], [8, rows * 8, rows * 8], [64, 128, 1, 1])
        taps = [(0, 0, (p - r) * P - p, 0, r * 64) for r in range(l.k)]
        no.add_conv_flat(plan, name, amap, taps, 64, self.w_dgrad[j], gin.ptr(n0, padded=True), l.cin,
                         rows, l.cin, b_mn=False)
      elif j in self.dY8:
        # MXFP8 dgrad: block-scaled dY x W^T (K = output channels), same flat zero-ringed domain
        d8, dsf = self.dY8[j], self.dYsf[j]
        taps = [(0, 0, (p - r) * P + (p - s), 0, (r * l.k + s) * l.cout)
                for r in range(l.k) for s in range(l.k)]
        no.add_conv_flat_mx(plan, name + ".mxfp8", d8.tmap_flat(n0, nimg), dsf.tmap_flat(n0, nimg, cbox=16, swizzle=0),
                            taps, l.cout, self.st.kernel_fp8_t(l), self.st.w_sft[self.st.fp8_slot(l)],
                            gin.ptr(n0, padded=True), l.cin, rows, l.cin,
                            halo_maps=lambda rb: (d8.tmap_flat(n0, nimg, rows_box=rb),
                                                  dsf.tmap_flat(n0, nimg, rows_box=rb, cbox=16, swizzle=0)))
      else:
        taps = [(0, 0, (p - r) * P + (p - s), 0, (r * l.k + s) * l.cin)
                for r in range(l.k) for s in range(l.k)]
        no.add_conv_flat(plan, name, dY.tmap_flat(n0, nimg), taps, l.cout, w2, gin.ptr(n0, padded=True),
                         l.cin, rows, l.cin, b_mn=True,
                         halo_map=lambda rows_box: dY.tmap_flat(n0, nimg, rows_box=rows_box))
      return (gin.view(n0), p)
    pb = self._pad_before(j)
    if l.stride == 2:
      bw, bh = nv.tile_shape(ho, wo)
      amap = dY.tmap_interior(bw, bh, n0, nimg)
      if is_small_cin(l):
        tgt, t0 = input_target
        view = (tgt.ptr(t0), tgt.sn, 2 * tgt.sh, 2 * tgt.sw)
        phases = [(no.taps_convT_phase(a, b, l.cout, l.k, pb), a * tgt.sh + b * tgt.sw)
                  for a in (0, 1) for b in (0, 1)]
        no.add_conv(plan, name, [amap], [], l.cout, self.w_dgrad[j], view, ho, wo, nimg, l.cin,
                    b_mn=False, phases=phases)
        return (tgt.view(t0), 0)
      gin = self.gin[j]
      view = (gin.ptr(n0), gin.sn, 2 * gin.sh, 2 * gin.sw)
      phases = [(no.taps_convT_phase(a, b, l.cin, l.k, pb), a * gin.sh + b * gin.sw)
                for a in (0, 1) for b in (0, 1)]
      no.add_conv(plan, name, [amap], [], l.cout, w2, view, ho, wo, nimg, l.cin, b_mn=True, phases=phases)
      return (gin.view(n0), 0)
    # stride-1 TF-SAME conv: dX[q] = sum dY[q + pb - (r, s)] W[r, s]
    gin = self.gin[j]
    bw, bh = nv.tile_shape(hi, wi)
    halo = None
    if nv.pick_block_n(l.cin) == 256 and hi % 16 == 0 and wi % 8 == 0:
      bw, bh = 8, 16
      halo = lambda w_, h_: dY.tmap_interior(w_, h_, n0, nimg)
    taps = [(0, 0, pb - s, pb - r, (r * l.k + s) * l.cin) for r in range(l.k) for s in range(l.k)]
    no.add_conv(plan, name, [dY.tmap_interior(bw, bh, n0, nimg)], taps, l.cout, w2, gin.view(n0), hi, wi,
                nimg, l.cin, b_mn=True, tile=(bw, bh), halo_map=halo)
    return (gin.view(n0), 0)
