"""Dry-run of the native step planner on CPU against a validating fake extension."""
import pytest
import torch

from tests import fake_native


class A:
  def __init__(self, nres):
    self.num_residual_blocks = nres


@pytest.mark.parametrize("size,b,nres", [(128, 2, 1), (256, 1, 9), (512, 1, 2), (96, 1, 1), ((192, 320), 1, 2),
                                         ((40, 72), 2, 1)])
def test_step_plan_builds_and_is_well_formed(monkeypatch, size, b, nres):
  fake_native.install(monkeypatch)
  monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
  from tf2_cyclegan_b200.engine.cyclegan import build_specs
  from tf2_cyclegan_b200.engine.native_backend import NETS, NativeBackend
  from tf2_cyclegan_b200.parallel.strategy import Strategy
  st = Strategy(device="cpu")
  be = NativeBackend(build_specs(A(nres)), b, st, seed=1, image_size=size, local_batch=b)
  sp = be.plan_for(b)
  names = sp.names
  assert sp.marks["cycle_end"] < sp.marks["test_end"] < sp.marks["G_bwd_end"] < sp.marks["end"]
  # six generator forwards = 2 batched + 2 cycle passes per generator pair
  assert sum(1 for n in names if n.startswith("G.stem.fwd")) == 2
  assert sum(1 for n in names if n.startswith("X.conv0.fwd")) == 1      # fakes forwarded once
  # gradient routing: no dgrad into a stem, adversarial pass has no wgrad
  assert not any("stem.dgrad" in n for n in names)
  adv = [n for n in names if n[0] in "XY" and n.endswith(f"[0:{b}]") or f"[0:{b}]." in n and n[0] in "XY"]
  assert adv and not any("wgrad" in n for n in adv) and any("conv0.dgrad" in n for n in adv)
  assert any(n.startswith("sync.fork") for n in names) and any(n.startswith("sync.join") for n in names)
  for n in NETS:
    assert f"opt.{n}.adam" in names and f"opt.{n}.step" in names
  # tap sharing: every residual 3x3 conv (forward: 8x16 boxed halo; dgrad: flat halo) qualifies
  kinds = {o[1]: o[0] for o in sp.plan.ops}
  if isinstance(size, int) and size >= 128:   # (other shapes: only that the plan builds and validates)
    res = [n for n in names if ".res" in n and (".fwd" in n or (".dgrad" in n and size <= 256))]
    assert res and all(kinds[n] == "conv_gemm_ts" for n in res), [n for n in res if kinds[n] != "conv_gemm_ts"][:4]
    if size > 256:   # flat halo of a 130-wide padded row (2 x 200 rows) exceeds the A stage: per-tap kernel
      assert all(kinds[n] == "conv_gemm" for n in names if ".res" in n and ".dgrad" in n)
  assert kinds[next(n for n in names if n.startswith("G.stem.fwd"))] == "conv_gemm"
  # every conv layer of every network gets exactly one wgrad group in the full backward
  full = names[sp.marks["test_end"]:]
  for l in be.specs["G"].layers:
    assert any(n.startswith(f"G.{l.name}.wgrad") for n in full), l.name
  for l in be.specs["X"].layers:
    assert any(n.startswith(f"X.{l.name}.wgrad") or n.startswith(f"X.{l.name}.bwd") for n in full), l.name


def test_flat_layout_roundtrip_and_mask():
  from tf2_cyclegan_b200.engine.native_params import FlatLayout
  from tf2_cyclegan_b200.models import oracle_net
  from tf2_cyclegan_b200.models.spec import discriminator_spec, generator_spec
  for spec in (generator_spec(num_residual_blocks=1), discriminator_spec()):
    lay = FlatLayout(spec)
    params = oracle_net.init_params(spec, torch.Generator().manual_seed(0))
    flat = lay.pack(params)
    back = lay.unpack(flat)
    for k, v in params.items():
      assert torch.equal(back[k], v), k
    mask = lay.mask()
    assert int(mask.sum()) == spec.num_params()
    assert (flat[mask == 0] == 0).all()
    # index tensor addresses the same elements
    l0 = spec.layers[0]
    idx = lay.index_tensor(l0)
    assert torch.equal(flat[idx], params[f"layer_with_weights-{l0.lw_index}/kernel"])


def test_fp8_plan_builds(monkeypatch):
  fake_native.install(monkeypatch)
  monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
  from tf2_cyclegan_b200.engine.cyclegan import build_specs
  from tf2_cyclegan_b200.engine.native_backend import NativeBackend
  from tf2_cyclegan_b200.parallel.strategy import Strategy
  be = NativeBackend(build_specs(A(2)), 2, Strategy(device="cpu"), seed=1, image_size=128, local_batch=2,
                     dtype="fp8")
  sp = be.plan_for(2)
  fp8_convs = [n for n in sp.names if n.endswith(".fp8")]
  # 2 generators x 2 residual blocks x 2 convs x 2 forward passes
  assert len(fp8_convs) == 16 and all(".res" in n for n in fp8_convs)
  assert "fp8.G.w_quant" in sp.names and "fp8.F.scales" in sp.names
  # backward stays bf16: same wgrad / dgrad op names as the bf16 plan
  assert any(n.startswith("G.res0a.wgrad") for n in sp.names)


def test_mxfp8_plan_builds(monkeypatch):
  """--dtype mxfp8: block-scaled convs + scale-record buffers + per-layer weight quantisation ops."""
  fake_native.install(monkeypatch)
  monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
  from tf2_cyclegan_b200.engine.cyclegan import build_specs
  from tf2_cyclegan_b200.engine.native_backend import NativeBackend
  from tf2_cyclegan_b200.parallel.strategy import Strategy
  be = NativeBackend(build_specs(A(2)), 2, Strategy(device="cpu"), seed=1, image_size=128, local_batch=2,
                     dtype="mxfp8")
  sp = be.plan_for(2)
  mx = [n for n in sp.names if n.endswith(".mxfp8") and ".fwd" in n]
  assert len(mx) == 16 and all(".res" in n for n in mx)
  mxd = [n for n in sp.names if n.endswith(".mxfp8") and ".dgrad" in n]
  bfd = [n for n in sp.names if ".res" in n and ".dgrad" in n and not n.endswith(".mxfp8")]
  assert len(mxd) > 0 and not bfd            # every residual-block dgrad runs block-scaled too
  assert sum(n.endswith(".mx_w_quant") for n in sp.names) == 8      # 2 generators x 4 residual convs
  assert sum(n.endswith(".mx_wt_quant") for n in sp.names) == 8     # ... and their dgrad operands W^T
  assert not any(n.endswith(".scales") or n.endswith(".w_amax") for n in sp.names)   # no scaling state
  ex = sp.ex["G"]
  assert set(ex.sf8) == set(ex.inp8) and all(b.c == 16 and b.halo == 1 for b in ex.sf8.values())
  assert set(ex.dY8) == set(ex.dYsf) == set(ex.inp8) and all(b.halo == 1 for b in ex.dY8.values())
  st = be.state["G"]
  assert [t.numel() for t in st.w_sf] == [256 * 9 * 256 // 32] * 4
  assert be.extra_state() == {}
  assert any(n.startswith("G.res0a.wgrad") for n in sp.names)


def test_mxfp8_op_builders_reject_bad_shapes(monkeypatch):
  fake_native.install(monkeypatch)
  import pytest
  from tf2_cyclegan_b200.engine import native_ops as no
  from tf2_cyclegan_b200.ops import native as nv
  plan = nv.load().Plan()
  w8 = torch.zeros(256, 9 * 256, dtype=torch.uint8)
  sfb = torch.zeros(256 * 9 * 256 // 32, dtype=torch.uint8)
  view = (1 << 20, 0, 0, 256)
  with pytest.raises(ValueError):      # 64 input channels: not a multiple of the 128-channel K block
    no.add_conv_mx(plan, "x", b"", b"", no.taps_conv(3, 64), 64, w8, sfb, view, 64, 64, 1, 256)
  with pytest.raises(ValueError):      # scale buffer of the wrong size
    no.add_conv_mx(plan, "x", b"", b"", no.taps_conv(3, 256), 256, w8, sfb[:-32], view, 64, 64, 1, 256)
  with pytest.raises(ValueError):      # 192 output channels: not a multiple of the 128-wide tile
    no.add_conv_flat_mx(plan, "x", b"", b"", [(0, 0, 0, 0, 0)], 256, w8, sfb, 1 << 20, 256, 4096, 192)


def test_allreduce_buckets_and_ownership(monkeypatch):
  """Bucket plan of the fused all-reduce (CPU logic only): buckets tile the kernel part of the flat
  buffer in backward order, the vector tail is one-shot (replicated), and the per-bucket rank shards
  (+ rank 0 for the tail) give every element exactly one owner in a checkpoint gather."""
  fake_native.install(monkeypatch)
  monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
  from tf2_cyclegan_b200.engine.cyclegan import build_specs
  from tf2_cyclegan_b200.engine.native_backend import NETS, NativeBackend
  from tf2_cyclegan_b200.parallel.fused_allreduce import SymmetricBuffers
  from tf2_cyclegan_b200.parallel.strategy import Strategy
  be = NativeBackend(build_specs(A(9)), 1, Strategy(device="cpu"), seed=1, image_size=128, local_batch=1)
  for n in NETS:
    bs = be.plan_buckets(n)
    st = be.state[n]
    vec_lo = st.layout.vec_lo
    # every vector parameter (bias / gamma / beta) lives in the tail, every kernel below it
    for seg in st.layout.segments:
      assert (seg.offset >= vec_lo) == (seg.kind == "vector"), seg.name
    assert 0 < st.n - vec_lo <= 32 * 1024 and vec_lo % 8 == 0     # <= 128 KB: the one-shot form
    assert len(bs) == (4 if n in "GF" else 2)
    assert bs[0][2] == vec_lo and bs[-1][0] == 0 and bs[-1][1] == 0
    for (j0, lo0, hi0), (j1, lo1, hi1) in zip(bs, bs[1:]):
      assert j1 < j0 and hi1 == lo0 and lo0 % 8 == 0           # descending layers, contiguous
    if n in "GF":
      assert max(hi - lo for _, lo, hi in bs) < 0.4 * st.n      # reasonably balanced (layer granularity)
    assert bs[-1][2] - bs[-1][1] < 0.3 * st.n                   # the exposed (last) bucket is small
    for world in (2, 3, 8):
      owners = torch.zeros(st.n, dtype=torch.int32)
      for rank in range(world):
        sb = SymmetricBuffers.__new__(SymmetricBuffers)
        sb.rank, sb.world, sb.buckets, sb.vec_lo = rank, world, {}, {}
        sb.set_buckets(n, sorted((lo, hi) for _, lo, hi in bs), vec_lo=vec_lo)
        for lo, hi in sb.owned_ranges(n, st.n):
          owners[lo:hi] += 1
      assert int(owners.min()) == 1 and int(owners.max()) == 1


def test_bucketed_allreduce_plan_order(monkeypatch):
  """Fused data-parallel plan (stub peer memory): every cross-GPU launch sits on the dedicated
  collective lane in ONE global order (comm_order: generators' buckets interleaved, then the
  discriminators'), each after events recorded on the main AND the wgrad lane once every op that
  writes the bucket's gradients has been issued; a network's last launch carries the vector tail
  (and, for SCALAR_NET, the loss scalars) and is followed by the per-step tail."""
  fake_native.install(monkeypatch)
  monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
  from tf2_cyclegan_b200.engine.cyclegan import build_specs
  from tf2_cyclegan_b200.engine.native_backend import COMM_LANE, NETS, SCALAR_NET, NativeBackend
  from tf2_cyclegan_b200.parallel.strategy import Strategy
  be = NativeBackend(build_specs(A(3)), 2, Strategy(device="cpu"), seed=1, image_size=128, local_batch=2)
  flags = {}

  class Symm:
    def emit_allreduce_adam(self, plan, n, st, bucket=None, vectors=None, scalars=False, exposed=False, **kw):
      assert set(kw) == {"lr", "b1", "b2", "eps"}
      flags[(n, bucket)] = (vectors, scalars, exposed)
      plan.add_memset(f"opt.{n}.AR.b{bucket}", st.g.data_ptr(), 16)

  be.fused, be.symm = True, Symm()
  for n in NETS:
    be.buckets[n] = be.plan_buckets(n)
    be.bucket_index[n] = list(range(len(be.buckets[n])))
  lanes = []
  import tests.fake_native as fn
  plan_cls = next(v for v in vars(fn).values() if isinstance(v, type) and hasattr(v, "_add"))
  orig = plan_cls._add

  def _add(self, kind, name, *args):
    lanes.append(getattr(self, "lane", 0))
    return orig(self, kind, name, *args)

  monkeypatch.setattr(plan_cls, "_add", _add)
  sp = be.plan_for(2)
  names = sp.names
  assert len(lanes) == len(names)
  pos = {nm: i for i, nm in enumerate(names)}
  # one global order, all on the collective lane
  ars = [i for i, nm in enumerate(names) if ".AR.b" in nm]
  assert [names[i] for i in ars] == [f"opt.{n}.AR.b{k}" for n, k in be.comm_order()]
  assert be.comm_order()[:4] == [("G", 0), ("F", 0), ("G", 1), ("F", 1)] and be.comm_order()[-1][0] == SCALAR_NET
  assert all(lanes[i] == COMM_LANE for i in ars)
  ops = sp.plan.ops
  for n, main_lane, wg_lane in (("G", 0, 2), ("Y", 0, 2), ("F", 1, 3), ("X", 1, 3)):
    layers = be.specs[n].layers
    nb = len(be.buckets[n])
    for k in range(nb):
      i = pos[f"opt.{n}.AR.b{k}"]
      last = k == nb - 1
      assert flags[(n, k)] == (last, last and n == SCALAR_NET, last)
      first_layer = be.buckets[n][k][0]
      last_layer = be.buckets[n][k - 1][0] - 1 if k else len(layers) - 1
      waits = [j for j in range(i) if names[j].startswith(f"sync.wait.{n}.bucket{k}")]
      recs = [ops[j][2][1] for j in waits]                     # the record op each wait refers to
      rec_lane = {r: ops[r][2][0] for r in recs}            # (lane,) of a record op
      assert sorted(rec_lane.values()) == [main_lane, wg_lane] and all(ops[j][2][0] == COMM_LANE for j in waits)
      for j in range(first_layer, last_layer + 1):
        grads = [q for q, nm in enumerate(names)
                 if nm.startswith(f"{n}.{layers[j].name}.") and ("wgrad" in nm or ".bwd[" in nm or "in_bwd_apply" in nm
                                                                  or "act_bwd" in nm)
                 and q > sp.marks["test_end"] and f"[0:{2 * (3 if n in 'GF' else 2)}]" in nm]
        assert grads, (n, k, layers[j].name)
        for q in grads:     # issued before the record op of its own lane
          assert q < next(r for r in recs if rec_lane[r] == lanes[q]), (n, k, names[q])
    tail = pos[f"opt.{n}.step"]
    assert tail > pos[f"opt.{n}.AR.b{nb - 1}"] and lanes[tail] == COMM_LANE


def test_plans_of_smaller_batches_share_the_arenas(monkeypatch):
  """The partial last batch of an epoch (or the b = 1 plot pass) must not double the activation
  memory: its plan uses the first images of the arenas allocated for the full batch."""
  fake_native.install(monkeypatch)
  monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
  from tf2_cyclegan_b200.engine.cyclegan import build_specs
  from tf2_cyclegan_b200.engine.native_backend import NativeBackend
  from tf2_cyclegan_b200.parallel.strategy import Strategy
  be = NativeBackend(build_specs(A(1)), 4, Strategy(device="cpu"), seed=1, image_size=64, local_batch=4)
  big, small = be.plan_for(4), be.plan_for(1)
  n_pool = len(be.buffer_pool)
  for n in "GFXY":
    eb, es = big.ex[n], small.ex[n]
    for bb, bs in list(zip(eb.inp, es.inp)) + [(a, b) for a, b in zip(eb.raw, es.raw) if a is not None] + \
                  [(a, b) for a, b in zip(eb.dY, es.dY) if a is not None] + [(eb.gout_final, es.gout_final)]:
      assert bs.storage.data_ptr() == bb.storage.data_ptr() and bs.n < bb.n and bs.sn == bb.sn
  be.plan_for(2)
  assert len(be.buffer_pool) == n_pool          # nothing new was allocated for b = 2
  bigger = be.plan_for(6)                        # a LARGER batch gets fresh arenas; older plans keep theirs
  assert bigger.ex["G"].inp[3].storage.data_ptr() != big.ex["G"].inp[3].storage.data_ptr()
