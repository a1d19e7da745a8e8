"""CPU stand-in for the native extension: records plan ops and validates every
TMA descriptor / launch geometry the planner produces, so the (GPU-only)
planning logic is exercised on the GPU-less pod."""
from __future__ import annotations

import struct


class FakePlan:

  def __init__(self):
    self.ops = []

  def size(self):
    return len(self.ops)

  def names(self):
    return [o[1] for o in self.ops]

  def run(self, *a, **k):
    raise RuntimeError("fake plan cannot run")

  def _add(self, kind, name, *args):
    self.ops.append((kind, name, args))
    return len(self.ops) - 1

  def add_conv_gemm(self, name, a_maps, b_map, taps, cin_blocks, tiles_w, tiles_h, bw, bh, w_limit,
                    out, sn, sh, sw, cout, bias, act, stats, block_n, b_mn, num_images, phases, out_scale=0, fp8=False,
                    ts_map=b"", ts=(), h_limit=0):
    if ts:   # tap sharing: one halo tile per channel block, taps = row shifts inside it
      nbox, box_step, box_bytes, pitch, odw, odh, sbo = ts
      td = decode(ts_map)
      assert block_n == 256 and not fp8 and not phases and len(a_maps) == 1
      assert td["box"][0] == 64 and td["box"][1] * td["box"][2] * 128 == box_bytes, (name, td["box"], box_bytes)
      assert nbox * box_bytes <= 34 * 1024 and (nbox == 1 or box_bytes % 1024 == 0)
      for t in taps:
        shift = (t[3] - odh) * pitch + (t[2] - odw)
        assert t[0] == 0 and t[1] == 0 and shift >= 0 and shift * 128 + 15 * sbo + 1024 <= nbox * box_bytes, (name, t)
    assert bw * bh == 128 and 1 <= len(a_maps) <= 4 and 1 <= len(taps) <= 52 and len(phases) <= 4
    if phases:
      assert sum(c for _, c, _ in phases) == len(taps) and all(c >= 1 for _, c, _ in phases)
    assert cin_blocks >= 1 and tiles_w >= 1 and tiles_h >= 1 and num_images >= 1
    assert block_n in (16, 32, 64, 128, 256) and (not b_mn or block_n >= 64)
    for t in taps:
      assert 0 <= t[0] < len(a_maps)
    kb = 128 if fp8 else 64
    for m in a_maps:
      d = decode(m)
      assert d["box"][:3] == [kb, bw, bh], (name, d["box"], bw, bh)
    bd = decode(b_map)
    if fp8:
      assert block_n == 256 and not b_mn and out_scale != 0 and bd["box"] == [128, 128]
      return self._add("conv_gemm_fp8", name, len(taps), cout, block_n, b_mn, num_images)
    assert bd["box"] == ([64, 64] if b_mn else [64, block_n // 2 if block_n >= 128 else block_n]), (name, bd["box"])
    if not b_mn:
      assert bd["dims"][1] >= min(cout, block_n), (name, bd["dims"], cout)
      for t in taps:
        assert t[4] + cin_blocks * 64 <= bd["dims"][0], (name, t, bd["dims"])
    else:
      assert bd["dims"][1] >= cin_blocks * 64
      for t in taps:
        assert t[4] + cout <= bd["dims"][0] + 63, (name, t, bd["dims"], cout)
    assert out % 16 == 0 or cout < 8
    return self._add("conv_gemm_ts" if ts else "conv_gemm", name, len(taps), cout, block_n, b_mn, num_images)

  def add_conv_gemm_mx(self, name, a_map, sfa_map, b_map, sfb, taps, cin_blocks, tiles_w, tiles_h, bw, bh, w_limit,
                       h_limit, out, sn, sh, sw, cout, bias, act, stats, num_images, ts_map=b"", sfts_map=b"", ts=()):
    assert bw * bh == 128 and 1 <= len(taps) <= 52 and 1 <= cin_blocks <= 4 and cout % 128 == 0
    if ts:   # tap sharing: activation halo box + scale-record halo box over the same pixels
      nbox, box_step, box_bytes, pitch, odw, odh, sbo = ts
      td, sd_ = decode(ts_map), decode(sfts_map)
      assert td["box"][0] == 128 and td["box"][1] * td["box"][2] * 128 == box_bytes, (name, td["box"])
      assert sd_["box"] == [16] + td["box"][1:] and sd_["dims"][1:] == td["dims"][1:], (name, sd_, td)
      assert nbox * box_bytes <= 34 * 1024 and (nbox == 1 or box_bytes % 1024 == 0)
      for t in taps:
        shift = (t[3] - odh) * pitch + (t[2] - odw)
        assert shift >= 0 and shift * 128 + 15 * sbo + 1024 <= nbox * box_bytes, (name, t)
    assert decode(a_map)["box"] == [128, bw, bh, 1], (name, decode(a_map)["box"])
    sd = decode(sfa_map)
    assert sd["box"] == [16, bw, bh, 1] and sd["dims"][0] == 16, (name, sd)
    assert sd["dims"][1:] == decode(a_map)["dims"][1:], "scale records must cover the same pixels"
    bd = decode(b_map)
    assert bd["box"] == [128, 64] and bd["dims"][1] >= cout
    for t in taps:
      assert t[0] == 0 and t[4] % (cin_blocks * 128) == 0 and t[4] + cin_blocks * 128 <= bd["dims"][0], (name, t)
    assert sfb % 16 == 0 and out % 16 == 0 and tiles_w >= 1 and tiles_h >= 1 and num_images >= 1
    return self._add("conv_gemm_mx", name, len(taps), cout, num_images)

  def add_wgrad(self, name, dy_maps, a_maps, taps, ktw, kth, kbw, kbh, nimg, cin, cout, dw, ldw, block_n, ksplit):
    assert kbw * kbh == 64 and ksplit != 0 and block_n in (64, 128, 256)
    if ksplit < 0:   # multi-tap kernel
      assert cin == 64 and 2 <= len(taps) <= 7 and len(dy_maps) == 1
    for m in dy_maps:
      assert decode(m)["box"][:3] == [64, kbw, kbh]
    for t in taps:
      assert 0 <= t[1] < len(dy_maps) and 0 <= t[0] < len(a_maps)
    for m in a_maps:
      assert decode(m)["box"][:3] == [64, kbw, kbh]
    for t in taps:
      assert t[4] + cin <= ldw, (name, t, cin, ldw)
    assert dw % 16 == 0 and ldw % 4 == 0
    return self._add("wgrad", name, len(taps), cin, cout)

  def set_lane(self, lane):
    self.lane = lane

  def add_fork(self, name):
    return self._add("fork", name)

  def add_join(self, name):
    return self._add("join", name)

  def add_event_record(self, name, lane):
    assert 0 <= lane < 5
    return self._add("record", name, lane)

  def add_event_wait(self, name, lane, ref):
    assert 0 <= lane < 5 and self.ops[ref][0] == "record"
    return self._add("wait", name, lane, ref)

  def set_nvtx(self, on):
    self.nvtx = bool(on)

  def __getattr__(self, item):
    if item.startswith("add_"):
      def f(name, *args):
        return self._add(item[4:], name, *args)
      return f
    raise AttributeError(item)


_MAPS = {}


def decode(b: bytes):
  return _MAPS[b]


class FakeModule:
  MAX_TAPS = 52
  Plan = FakePlan
  _counter = 0

  @staticmethod
  def encode_tmap(ptr, dtype, dims, strides_bytes, box, swizzle=3):
    rank = len(dims)
    assert 1 <= rank <= 5 and len(box) == rank and len(strides_bytes) == rank - 1
    assert ptr % 16 == 0, "TMA base must be 16-byte aligned"
    for s in strides_bytes:
      assert s % 16 == 0 and 0 < s < (1 << 40), f"bad stride {s}"
    for d in dims:
      assert 1 <= d <= (1 << 32)
    for b_ in box:
      assert 1 <= b_ <= 256
    esz = {0: 2, 1: 4, 2: 1}[dtype]
    elem = esz
    assert box[0] * esz == 128 or swizzle != 3, "SWIZZLE_128B needs a 128-byte inner box"
    FakeModule._counter += 1
    key = struct.pack("<Q", FakeModule._counter) + b"\0" * 120
    _MAPS[key] = {"ptr": ptr, "dims": list(dims), "strides": list(strides_bytes), "box": list(box)}
    return key


def install(monkeypatch):
  from tf2_cyclegan_b200.ops import native as nv
  monkeypatch.setattr(nv, "_mod", FakeModule)
  monkeypatch.setattr(nv, "available", lambda: True)
  return FakeModule
