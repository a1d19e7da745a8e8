"""sm_100a kernels vs the PyTorch fp32 oracle (SURVEY §4 'kernel (GPU)').

Inputs / weights are rounded to bf16 first so the only differences are fp32
accumulation order and the bf16 rounding of stored outputs.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _mods():
  from tf2_cyclegan_b200.engine import native_ops as no
  from tf2_cyclegan_b200.ops import native as nv
  from tf2_cyclegan_b200.ops import oracle as O
  return nv, no, O


def bf(t):
  return t.to(torch.bfloat16).float()


def rel_err(a, b):
  return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-6)).item()


def run(plan):
  plan.run(0, -1, torch.cuda.current_stream().cuda_stream, True)
  torch.cuda.synchronize()


def fill_reflect(buf, x):
  """reference fill of a haloed buffer (python side)."""
  nv, no, O = _mods()
  buf.t.copy_(O.reflect_pad(x, buf.halo).to(torch.bfloat16))


@pytest.mark.parametrize("n,h,c,cout", [(2, 64, 256, 256), (1, 16, 64, 128), (3, 32, 128, 512)])
def test_conv3x3_reflect_box(n, h, c, cout):
  nv, no, O = _mods()
  C = nv.load()
  torch.manual_seed(0)
  x = bf(torch.randn(n, h, h, c, device="cuda"))
  k = bf(torch.randn(3, 3, c, cout, device="cuda") * 0.05)
  xin = nv.Buf(n, h, h, c, halo=1)
  fill_reflect(xin, x)
  w2 = no.pack_conv_weight(k).to(torch.bfloat16)
  out = nv.Buf(n, h, h, cout)
  stats = torch.zeros(n, cout, 2, device="cuda")
  plan = C.Plan()
  bw, bh = nv.tile_shape(h, h)
  no.add_conv(plan, "conv", [xin.tmap_padded(bw, bh)], no.taps_conv(3, c), c, w2, out.view(), h, h,
              n, cout, stats=stats.data_ptr())
  run(plan)
  ref = O.conv2d(O.reflect_pad(x, 1), k, 1, "valid")
  assert rel_err(out.t, ref) < 1e-2
  got = out.t.float()
  assert torch.allclose(stats[..., 0], got.sum((1, 2)), rtol=1e-3, atol=1e-2 * h)
  assert torch.allclose(stats[..., 1], (got * got).sum((1, 2)), rtol=1e-3, atol=1e-2 * h)


@pytest.mark.parametrize("k,pad_before,cin,cout,h", [(3, 0, 64, 128, 32), (4, 1, 64, 128, 32),
                                                     (4, 1, 128, 256, 64)])
def test_conv_stride2_parity(k, pad_before, cin, cout, h):
  nv, no, O = _mods()
  C = nv.load()
  torch.manual_seed(1)
  n = 2
  x = bf(torch.randn(n, h, h, cin, device="cuda"))
  kern = bf(torch.randn(k, k, cin, cout, device="cuda") * 0.05)
  xin = nv.Buf(n, h, h, cin)
  xin.t.copy_(x.to(torch.bfloat16))
  oh = h // 2
  bw, bh = nv.tile_shape(oh, oh)
  maps = [xin.tmap_parity(a, b, bw, bh) for a in (0, 1) for b in (0, 1)]
  w2 = no.pack_conv_weight(kern).to(torch.bfloat16)
  out = nv.Buf(n, oh, oh, cout)
  plan = C.Plan()
  no.add_conv(plan, "s2", maps, no.taps_conv_s2(k, cin, pad_before), cin, w2, out.view(), oh, oh, n, cout)
  run(plan)
  ref = O.conv2d(x, kern, 2, "same")
  assert O.same_pad_amounts(h, k, 2)[0] == pad_before
  assert rel_err(out.t, ref) < 1e-2


@pytest.mark.parametrize("cout,bias,act", [(512, False, None), (1, True, None), (3, True, "tanh")])
def test_conv4x4_s1_same_and_small_n(cout, bias, act):
  nv, no, O = _mods()
  C = nv.load()
  torch.manual_seed(2)
  n, h, cin = 2, 32, 256
  x = bf(torch.randn(n, h, h, cin, device="cuda"))
  kern = bf(torch.randn(4, 4, cin, cout, device="cuda") * 0.03)
  b = torch.randn(cout, device="cuda") if bias else None
  xin = nv.Buf(n, h, h, cin)
  xin.t.copy_(x.to(torch.bfloat16))
  bw, bh = nv.tile_shape(h, h)
  w2 = no.pack_conv_weight(kern).to(torch.bfloat16)
  cpad = 8 if cout < 8 else cout
  out = nv.Buf(n, h, h, cpad)
  plan = C.Plan()
  no.add_conv(plan, "c4", [xin.tmap_interior(bw, bh)], no.taps_conv(4, cin, origin=-1), cin, w2,
              out.view(), h, h, n, cout, bias=no.fptr(b), act=nv.ACT[act])
  run(plan)
  ref = O.conv2d(x, kern, 1, "same", b)
  if act == "tanh":
    ref = torch.tanh(ref)
  assert rel_err(out.t[..., :cout], ref) < 1e-2
  if cpad > cout:
    assert out.t[..., cout:].abs().max() == 0


def test_conv_transpose_phases():
  nv, no, O = _mods()
  C = nv.load()
  torch.manual_seed(3)
  n, h, cin, cout = 2, 32, 128, 64
  x = bf(torch.randn(n, h, h, cin, device="cuda"))
  kern = bf(torch.randn(3, 3, cout, cin, device="cuda") * 0.05)   # [kh, kw, out, in]
  xin = nv.Buf(n, h, h, cin)
  xin.t.copy_(x.to(torch.bfloat16))
  w2 = no.pack_conv_weight(kern, transpose_conv=True).to(torch.bfloat16)
  out = nv.Buf(n, 2 * h, 2 * h, cout)
  stats = torch.zeros(n, cout, 2, device="cuda")
  bw, bh = nv.tile_shape(h, h)
  plan = C.Plan()
  view = (out.ptr(0), out.sn, 2 * out.sh, 2 * out.sw)
  phases = [(no.taps_convT_phase(a, b, cin), a * out.sh + b * out.sw) for a in (0, 1) for b in (0, 1)]
  no.add_conv(plan, "up", [xin.tmap_interior(bw, bh)], [], cin, w2, view, h, h, n, cout,
              stats=stats.data_ptr(), phases=phases)
  run(plan)
  ref = O.conv2d_transpose(x, kern, 2)
  assert rel_err(out.t, ref) < 1e-2
  assert torch.allclose(stats[..., 0], out.t.float().sum((1, 2)), rtol=1e-3, atol=0.5)


def test_dgrad_flat_mn_major_and_fold():
  """dgrad of the reflect-padded 3x3 conv in the zero-ringed flat domain, using the
  forward weight matrix as an MN-major B operand."""
  nv, no, O = _mods()
  C = nv.load()
  torch.manual_seed(4)
  n, h, c, cout = 2, 32, 128, 256
  x = bf(torch.randn(n, h, h, c, device="cuda")).requires_grad_(True)
  kern = bf(torch.randn(3, 3, c, cout, device="cuda") * 0.05)
  dy = bf(torch.randn(n, h, h, cout, device="cuda"))
  y = O.conv2d(O.reflect_pad(x, 1), kern, 1, "valid")
  (gx,) = torch.autograd.grad(y, x, dy)

  dyb = nv.Buf(n, h, h, cout, halo=1)           # zero ring
  dyb.interior().copy_(dy.to(torch.bfloat16))
  w2 = no.pack_conv_weight(kern).to(torch.bfloat16)
  P = h + 2
  rows = n * P * P
  dxpad = nv.Buf(n, h, h, c, halo=1)
  # dXpad[q] = sum_{r,s} dYz[q + (1 - r) * P + (1 - s)] W[r, s]
  taps = [(0, 0, (1 - r) * P + (1 - s), 0, (r * 3 + s) * c) for r in range(3) for s in range(3)]
  plan = C.Plan()
  no.add_conv_flat(plan, "dgrad", dyb.tmap_flat(), taps, cout, w2, dxpad.ptr(padded=True), c, rows, c,
                   b_mn=True)
  # fold the halo back with the IN-backward reduce kernel (no norm: dy = folded g)
  g = nv.Buf(n, h, h, c)
  plan.add_in_bwd_reduce("fold", dxpad.view(), 1, nv.NULL_VIEW, nv.NULL_VIEW, 0, 0, 0, 1.0, 1e-3, 0,
                         g.view(), 0, n, h, h, c)
  run(plan)
  assert rel_err(g.t, gx) < 1.5e-2


@pytest.mark.parametrize("n,h,cin,cout", [(2, 32, 256, 256), (2, 32, 128, 64)])
def test_wgrad_conv3x3(n, h, cin, cout):
  nv, no, O = _mods()
  C = nv.load()
  torch.manual_seed(5)
  x = bf(torch.randn(n, h, h, cin, device="cuda"))
  kern = bf(torch.randn(3, 3, cin, cout, device="cuda") * 0.05).requires_grad_(True)
  dy = bf(torch.randn(n, h, h, cout, device="cuda"))
  y = O.conv2d(O.reflect_pad(x, 1), kern, 1, "valid")
  (gk,) = torch.autograd.grad(y, kern, dy)

  xin = nv.Buf(n, h, h, cin, halo=1)
  fill_reflect(xin, x)
  dyb = nv.Buf(n, h, h, cout, halo=1)
  dyb.interior().copy_(dy.to(torch.bfloat16))
  dw = torch.zeros(max(cout, 16), 9 * cin, device="cuda")
  kbw, kbh = nv.tile_shape(h, h, 64)
  plan = C.Plan()
  no.add_wgrad(plan, "wgrad", dyb.tmap_interior(kbw, kbh), [xin.tmap_padded(kbw, kbh)],
               no.taps_conv(3, cin), h // kbw, h // kbh, kbw, kbh, n, cin, cout, dw.data_ptr(), 9 * cin)
  run(plan)
  got = no.unpack_conv_weight(dw, 3, cin, cout)
  assert rel_err(got, gk) < 1e-2


def test_overlapped_stride_stem():
  """7x7 stem conv (Cin = 3) as a GEMM: 8-channel pixels, one tap per kernel row, the 64-wide
  K block is an overlapping window of 8 consecutive pixels (TMA map with a 16-byte w stride)."""
  nv, no, O = _mods()
  C = nv.load()
  torch.manual_seed(6)
  n, h = 2, 64
  x = bf(torch.randn(n, h, h, 3, device="cuda"))
  kern = bf(torch.randn(7, 7, 3, 64, device="cuda") * 0.1)
  xin = nv.Buf(n, h, h, 8, halo=3, slack_rows=1)
  xp = O.reflect_pad(x, 3)
  xin.t[..., :3].copy_(xp.to(torch.bfloat16))
  # packed weights W2[co][r*64 + s*8 + c]
  w2 = torch.zeros(64, 7, 8, 8, device="cuda")
  w2[:, :, :7, :3] = kern.permute(3, 0, 1, 2)
  w2 = w2.reshape(64, 448).to(torch.bfloat16)
  bw, bh = nv.tile_shape(h, h)
  amap = nv.tmap(xin.ptr(padded=True), [64, h, xin.hp, n], [8, xin.sh, xin.sn], [64, bw, bh, 1])
  out = nv.Buf(n, h, h, 64)
  taps = [(0, 0, 0, r, r * 64) for r in range(7)]
  plan = C.Plan()
  no.add_conv(plan, "stem", [amap], taps, 64, w2, out.view(), h, h, n, 64)
  run(plan)
  ref = O.conv2d(xp, kern, 1, "valid")
  assert rel_err(out.t, ref) < 1e-2


def rel_l2(a, b):
  return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


@pytest.mark.parametrize("stream", [True, False])
@pytest.mark.parametrize("n,h,w,c,act,with_skip,with_g2", [
    (2, 16, 16, 64, "relu", True, False),
    (3, 64, 64, 256, "relu", False, False),     # residual-block 'a' shape: several chunks per image
    (3, 64, 64, 256, "none", True, True),       # residual-block 'b': skip input, second gradient
    (2, 12, 20, 128, "leaky", False, True),     # W not a power of two: stages straddle rows
    (1, 32, 32, 512, "leaky", False, False),    # PatchGAN conv3
])
def test_in_apply_and_backward_match_autograd(stream, n, h, w, c, act, with_skip, with_g2, request):
  """InstanceNorm apply (+activation, +residual, +reflect halo) and its two backward kernels vs
  autograd of the fp32 oracle — both the bulk-copy-ring kernels (default) and the direct ones."""
  nv, no, O = _mods()
  C = nv.load()
  C.set_ew_stream(stream)
  request.addfinalizer(lambda: C.set_ew_stream(True))
  torch.manual_seed(7)
  x = bf(torch.randn(n, h, w, c, device="cuda") * 2 + 0.5).requires_grad_(True)
  gamma = (torch.randn(c, device="cuda") * 0.5).requires_grad_(True)
  beta = (torch.randn(c, device="cuda") * 0.1).requires_grad_(True)
  skip = bf(torch.randn(n, h, w, c, device="cuda"))
  y = O.instance_norm(x, gamma, beta)
  y = torch.relu(y) if act == "relu" else (torch.nn.functional.leaky_relu(y, 0.2) if act == "leaky" else y)
  if with_skip:
    y = y + skip
  ypad = O.reflect_pad(y, 1)
  gpad = bf(torch.randn_like(ypad))
  g2 = bf(torch.randn(n, h, w, c, device="cuda")) if with_g2 else None
  gx, gg, gb = torch.autograd.grad([ypad] + ([y] if with_g2 else []), [x, gamma, beta],
                                   [gpad] + ([g2] if with_g2 else []))

  xb = nv.Buf(n, h, w, c); xb.t.copy_(x.detach().to(torch.bfloat16))
  sb = nv.Buf(n, h, w, c, halo=1); sb.interior().copy_(skip.to(torch.bfloat16))
  stats = torch.stack([x.detach().sum((1, 2)), (x.detach() ** 2).sum((1, 2))], -1).contiguous()
  dst = nv.Buf(n, h, w, c, halo=1)
  gp = nv.Buf(n, h, w, c, halo=1); gp.t.copy_(gpad.to(torch.bfloat16))
  g2b = nv.Buf(n, h, w, c)
  if with_g2:
    g2b.t.copy_(g2.to(torch.bfloat16))
  dyb = nv.Buf(n, h, w, c)
  dxb = nv.Buf(n, h, w, c, halo=1)
  bst = torch.zeros(n, c, 2, device="cuda")
  dgam = torch.zeros(c, device="cuda"); dbet = torch.zeros(c, device="cuda")
  inv = 1.0 / (h * w)
  plan = C.Plan()
  plan.add_in_apply("apply", xb.view(), stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), inv, 1e-3,
                    nv.ACT[act], sb.view() if with_skip else nv.NULL_VIEW, dst.view(), 1, n, h, w, c)
  plan.add_in_bwd_reduce("red", gp.view(), 1, g2b.view() if with_g2 else nv.NULL_VIEW, xb.view(), stats.data_ptr(),
                         gamma.data_ptr(), beta.data_ptr(), inv, 1e-3, nv.ACT[act], dyb.view(),
                         bst.data_ptr(), n, h, w, c)
  plan.add_in_bwd_apply("app", dyb.view(), xb.view(), stats.data_ptr(), bst.data_ptr(), gamma.data_ptr(),
                        inv, 1e-3, dxb.view(), dgam.data_ptr(), dbet.data_ptr(), n, h, w, c)
  run(plan)
  assert rel_err(dst.t, ypad) < 1e-2 and rel_l2(dst.t, ypad) < 5e-3
  assert rel_err(dxb.interior(), gx) < 2e-2 and rel_l2(dxb.interior(), gx) < 1e-2
  assert dxb.t[:, 0].abs().max() == 0 and dxb.t[:, :, 0].abs().max() == 0   # zero ring untouched
  assert rel_err(dgam, gg) < 2e-2 and rel_err(dbet, gb) < 2e-2
  assert rel_l2(dgam, gg) < 1e-2 and rel_l2(dbet, gb) < 1e-2


def test_losses_adam_imgprep():
  nv, no, O = _mods()
  C = nv.load()
  torch.manual_seed(8)
  n, h = 2, 32
  img = torch.rand(n, h, h, 3, device="cuda") * 2 - 1
  pred = nv.Buf(n, h, h, 8)
  pred.t[..., :3].copy_((torch.rand(n, h, h, 3, device="cuda") * 2 - 1).to(torch.bfloat16))
  g = nv.Buf(n, h, h, 8)
  slots = torch.zeros(4, device="cuda")
  refl = nv.Buf(n, h, h, 8, halo=3)
  plain = nv.Buf(n, h, h, 8, halo=1)
  d = nv.Buf(n, 8, 8, 8)
  d.t[..., 0].copy_(torch.randn(n, 8, 8, device="cuda").to(torch.bfloat16))
  gd = nv.Buf(n, 8, 8, 8)
  plan = C.Plan()
  gbs = 4
  scale = 10.0 / (h * h * 3 * gbs)
  plan.add_l1_loss("l1", pred.view(), img.data_ptr(), scale, scale, slots.data_ptr(),
                   slots[1:].data_ptr(), 1.0 / (h * h * 3 * gbs), g.view(), n, h, h)
  plan.add_mse_loss("mse", d.view(), 1.0, 1.0 / (64 * gbs), 2.0 / (64 * gbs), slots[2:].data_ptr(),
                    gd.view(), n, 8, 8)
  plan.add_img_prep("prep", img.data_ptr(), nv.NULL_VIEW, refl.view(), 3, plain.view(), n, h, h)
  # adam
  w = torch.randn(1024, device="cuda"); w0 = w.clone()
  gr = torch.randn(1024, device="cuda")
  m = torch.zeros(1024, device="cuda"); v = torch.zeros(1024, device="cuda")
  wb = torch.zeros(1024, device="cuda", dtype=torch.bfloat16)
  step = torch.zeros(1, device="cuda", dtype=torch.int32)
  for _ in range(2):
    plan.add_adam("adam", w.data_ptr(), m.data_ptr(), v.data_ptr(), gr.data_ptr(), wb.data_ptr(),
                  step.data_ptr(), 2e-4, 0.5, 0.9, 1e-7, 1.0, 1024)
    plan.add_step_increment("inc", step.data_ptr())
  run(plan)
  p3 = pred.t[..., :3].float()
  assert math.isclose(slots[0].item(), 10 * (img - p3).abs().flatten(1).mean(1).sum().item() / gbs, rel_tol=1e-3)
  assert math.isclose(slots[1].item(), (img - p3).abs().flatten(1).mean(1).sum().item() / gbs, rel_tol=1e-3)
  assert torch.allclose(g.t[..., :3].float(), -scale * torch.sign(img - p3), atol=1e-3 * scale, rtol=1e-2)
  dd = d.t[..., 0].float()
  assert math.isclose(slots[2].item(), ((1 - dd) ** 2).flatten(1).mean(1).sum().item() / gbs, rel_tol=1e-3)
  assert rel_err(gd.t[..., 0], 2.0 / (64 * gbs) * (dd - 1)) < 1e-2
  assert rel_err(refl.t[..., :3], O.reflect_pad(img, 3)) < 1e-2
  assert rel_err(plain.interior()[..., :3], img) < 1e-2 and plain.t[:, 0].abs().max() == 0
  from tf2_cyclegan_b200.ops.keras_adam import KerasAdam
  opt = KerasAdam()
  ref = {"w": w0.clone()}
  for _ in range(2):
    opt.apply_gradients(ref, {"w": gr})
  assert torch.allclose(w, ref["w"], atol=1e-6) and step.item() == 2
  assert torch.allclose(wb.float(), w, atol=1e-2)


def test_output_conv_row_gemm_and_gather():
  """7x7 reflect conv 64 -> 3 (+bias, tanh) as one GEMM tap per kernel row over the flat padded
  domain followed by the 7-tap horizontal gather."""
  nv, no, O = _mods()
  C = nv.load()
  torch.manual_seed(9)
  n, h, cin, cout, k = 2, 64, 64, 3, 7
  x = bf(torch.randn(n, h, h, cin, device="cuda"))
  kern = bf(torch.randn(k, k, cin, cout, device="cuda") * 0.05)
  bias = torch.randn(cout, device="cuda") * 0.1
  xin = nv.Buf(n, h, h, cin, halo=3)
  fill_reflect(xin, x)
  wq = torch.zeros(32, k, cin, device="cuda")
  for s in range(k):
    wq[s * 4:s * 4 + cout] = kern[:, s].permute(2, 0, 1)      # [co, r, ci]
  wq = wq.reshape(32, k * cin).to(torch.bfloat16)
  rows = n * xin.hp * xin.wp
  q = torch.zeros((rows + 256) * 32, dtype=torch.bfloat16, device="cuda")
  out = nv.Buf(n, h, h, 8)
  plan = C.Plan()
  taps = [(0, 0, r * xin.wp, 0, r * cin) for r in range(k)]
  no.add_conv_flat(plan, "rowgemm", xin.tmap_flat(), taps, cin, wq, q.data_ptr(), 32, rows, 32, block_n=32)
  plan.add_tap_gather("gather", q.data_ptr(), bias.data_ptr(), out.view(), nv.ACT["tanh"], k, xin.hp, xin.wp,
                      n, h, h)
  run(plan)
  ref = torch.tanh(O.conv2d(O.reflect_pad(x, 3), kern, 1, "valid", bias))
  assert (out.t[..., :cout].float() - ref).abs().max().item() < 3e-2
  assert out.t[..., cout:].abs().max() == 0


def test_fp8_conv_and_in_apply_quantisation():
  """E4M3 x E4M3 conv (kind::f8f6f4, cta_group::2) with epilogue de-scaling, fed by the FP8 side
  output of the IN-apply kernel; oracle = fp32 conv of the de-quantised operands."""
  nv, no, O = _mods()
  C = nv.load()
  torch.manual_seed(11)
  n, h, c, cout = 2, 64, 256, 256
  x = bf(torch.randn(n, h, h, c, device="cuda"))
  kern = bf(torch.randn(3, 3, c, cout, device="cuda") * 0.05)
  sa, sb = 32.0, 1024.0
  # producer: identity IN-apply (no stats) writing bf16 + E4M3 reflect-padded copies
  xb = nv.Buf(n, h, h, c); xb.t.copy_(x.to(torch.bfloat16))
  dst = nv.Buf(n, h, h, c, halo=1)
  dst8 = nv.Buf8(n, h, h, c, halo=1)
  scale = torch.tensor([sa], device="cuda"); amax = torch.zeros(1, device="cuda")
  oscale = torch.tensor([1.0 / (sa * sb)], device="cuda")
  w2 = no.pack_conv_weight(kern)
  w8 = (w2 * sb).to(torch.float8_e4m3fn).view(torch.uint8).contiguous()
  out = nv.Buf(n, h, h, cout)
  stats = torch.zeros(n, cout, 2, device="cuda")
  plan = C.Plan()
  plan.add_in_apply("q", xb.view(), 0, 0, 0, 1.0, 1e-3, 0, nv.NULL_VIEW, dst.view(), 1, n, h, h, c,
                    dst8.view(), scale.data_ptr(), amax.data_ptr())
  bw, bh = nv.tile_shape(h, h)
  no.add_conv(plan, "conv.fp8", [dst8.tmap_padded(bw, bh)], no.taps_conv(3, c), c, w8, out.view(), h, h, n,
              cout, stats=stats.data_ptr(), fp8=True, out_scale=oscale.data_ptr())
  run(plan)
  assert abs(amax.item() - x.abs().max().item()) < 1e-2
  xq = (O.reflect_pad(x, 1) * sa).to(torch.float8_e4m3fn).float() / sa
  got8 = dst8.t.view(torch.float8_e4m3fn).float() / sa
  assert (got8 - xq).abs().max().item() == 0
  wq = no.unpack_conv_weight((w2 * sb).to(torch.float8_e4m3fn).float() / sb, 3, c, cout)
  ref = O.conv2d(xq, wq, 1, "valid")
  assert rel_err(out.t, ref) < 1e-2
  full = O.conv2d(O.reflect_pad(x, 1), kern, 1, "valid")
  assert rel_err(out.t, full) < 8e-2          # quantisation error of E4M3 operands


@pytest.mark.parametrize("shared", [False, True])
def test_mxfp8_conv_and_quantisers(shared):
  """MXFP8 (E4M3 + E8M0 scale per 32 channels): the IN-apply twin writer and the weight quantiser must
  match the PyTorch reference bit for bit, and the block-scaled conv (kind::mxf8f6f4.block_scale,
  scales staged into TMEM) must equal an fp32 conv of the de-quantised operands. ``shared``: the
  tap-shared CTA-pair kernel (8 x 16 tiles, halo tiles of activations and scale records)."""
  nv, no, O = _mods()
  from tf2_cyclegan_b200.ops import mxfp8
  C = nv.load()
  torch.manual_seed(12)
  n, h, c, cout = 2, 64, 256, 256
  # block-dependent magnitudes so that neighbouring scales differ by many powers of two
  mag = torch.exp2(torch.randint(-6, 7, (n, h, h, c // 32), device="cuda").float()).repeat_interleave(32, dim=-1)
  x = bf(torch.randn(n, h, h, c, device="cuda") * mag)
  kmag = torch.exp2(torch.randint(-4, 5, (3, 3, c // 32, cout), device="cuda").float()).repeat_interleave(32, dim=2)
  kern = bf(torch.randn(3, 3, c, cout, device="cuda") * 0.05 * kmag)
  xb = nv.Buf(n, h, h, c); xb.t.copy_(x.to(torch.bfloat16))
  dst = nv.Buf(n, h, h, c, halo=1)
  dst8 = nv.Buf8(n, h, h, c, halo=1)
  sf8 = nv.Buf8(n, h, h, 16, halo=1)
  w2 = no.pack_conv_weight(kern).to(torch.bfloat16).contiguous()       # [cout, 9 * c]
  w8 = torch.zeros(cout, 9 * c, dtype=torch.uint8, device="cuda")
  wsf = torch.zeros(cout * 9 * c // 32, dtype=torch.uint8, device="cuda")
  out = nv.Buf(n, h, h, cout)
  stats = torch.zeros(n, cout, 2, device="cuda")
  plan = C.Plan()
  plan.add_in_apply("q", xb.view(), 0, 0, 0, 1.0, 1e-3, 0, nv.NULL_VIEW, dst.view(), 1, n, h, h, c,
                    dst8.view(), 0, 0, sf8.view())
  no.add_mx_weight_quant(plan, "wq", w2.data_ptr(), w8.data_ptr(), wsf.data_ptr(), cout, 9, c)
  bw, bh = (8, 16) if shared else nv.tile_shape(h, h)
  halo = (lambda w_, h_: (dst8.tmap_padded(w_, h_), sf8.tmap_padded(w_, h_, cbox=16, swizzle=0))) if shared else None
  no.add_conv_mx(plan, "conv.mx", dst8.tmap_padded(bw, bh), sf8.tmap_padded(bw, bh, cbox=16, swizzle=0),
                 no.taps_conv(3, c), c, w8, wsf, out.view(), h, h, n, cout, stats=stats.data_ptr(),
                 tile=(bw, bh), halo_maps=halo)
  run(plan)
  # activations: bytes + scale records (reflect halo included), bit-exact
  xp = O.reflect_pad(x, 1)
  q_ref, sb_ref = mxfp8.quantize(xp)
  assert torch.equal(dst8.t, q_ref)
  assert torch.equal(sf8.t[..., :c // 32], sb_ref) and (sf8.t[..., c // 32:] == 0).all()
  assert (dst.t.float() - xp).abs().max().item() == 0
  # weights
  wq_ref, wsb_ref = mxfp8.quantize(w2.float())
  assert torch.equal(w8, wq_ref)
  assert torch.equal(wsf, mxfp8.weight_scale_order(wsb_ref, 9, c))
  # conv vs the fp32 conv of the de-quantised operands
  xq = mxfp8.dequantize(q_ref, sb_ref)
  wq = no.unpack_conv_weight(mxfp8.dequantize(wq_ref, wsb_ref), 3, c, cout)
  ref = O.conv2d(xq, wq, 1, "valid")
  assert rel_err(out.t, ref) < 1e-2 and rel_l2(out.t, ref) < 5e-3
  s_ref = torch.stack([out.t.float().sum((1, 2)), (out.t.float() ** 2).sum((1, 2))], -1)
  assert rel_err(stats, s_ref) < 2e-3
  full = O.conv2d(xp, kern, 1, "valid")
  assert rel_l2(out.t, full) < 6e-2            # quantisation error of E4M3 operands with block scales


@pytest.mark.parametrize("shared", [False, True])
def test_mxfp8_flat_dgrad(shared):
  """MXFP8 dgrad of the reflect-padded 3x3 conv in the zero-ringed flat domain: block-scaled dY (K = output
  channels) x W^T from the strided weight quantiser. Oracle = autograd of the fp32 conv on the de-quantised
  operands; ``shared``: one 272-row halo tile (two boxes) of bytes AND scale records per channel block."""
  nv, no, O = _mods()
  from tf2_cyclegan_b200.ops import mxfp8
  C = nv.load()
  torch.manual_seed(13)
  n, h, c, cout = 2, 32, 256, 256
  kmag = torch.exp2(torch.randint(-4, 5, (3, 3, c, cout // 32), device="cuda").float()).repeat_interleave(32, dim=3)
  kern = bf(torch.randn(3, 3, c, cout, device="cuda") * 0.05 * kmag)
  mag = torch.exp2(torch.randint(-6, 7, (n, h, h, cout // 32), device="cuda").float()).repeat_interleave(32, dim=-1)
  dy = bf(torch.randn(n, h, h, cout, device="cuda") * mag)
  # operands: dY twin (zero ring) from the reference quantiser, W^T from the kernel under test
  q, sb = mxfp8.quantize(dy)
  d8 = nv.Buf8(n, h, h, cout, halo=1)
  dsf = nv.Buf8(n, h, h, 16, halo=1)
  d8.interior().copy_(q)
  dsf.interior()[..., :cout // 32].copy_(sb)
  w2 = no.pack_conv_weight(kern).to(torch.bfloat16).contiguous()          # [cout, 9 * c]
  w8t = torch.zeros(c, 9 * cout, dtype=torch.uint8, device="cuda")
  wsft = torch.zeros(c * 9 * cout // 32, dtype=torch.uint8, device="cuda")
  P = h + 2
  rows = n * P * P
  dxpad = nv.Buf(n, h, h, c, halo=1)
  taps = [(0, 0, (1 - r) * P + (1 - s), 0, (r * 3 + s) * cout) for r in range(3) for s in range(3)]
  plan = C.Plan()
  no.add_mx_weight_quant(plan, "wtq", w2.data_ptr(), w8t.data_ptr(), wsft.data_ptr(), cout, 9, c, transposed=True)
  halo = (lambda rb: (d8.tmap_flat(rows_box=rb), dsf.tmap_flat(rows_box=rb, cbox=16, swizzle=0))) if shared else None
  no.add_conv_flat_mx(plan, "dgrad.mx", d8.tmap_flat(), dsf.tmap_flat(cbox=16, swizzle=0), taps, cout, w8t, wsft,
                      dxpad.ptr(padded=True), c, rows, c, halo_maps=halo)
  g = nv.Buf(n, h, h, c)
  plan.add_in_bwd_reduce("fold", dxpad.view(), 1, nv.NULL_VIEW, nv.NULL_VIEW, 0, 0, 0, 1.0, 1e-3, 0,
                         g.view(), 0, n, h, h, c)
  run(plan)
  # W^T[ci][tap * cout + co] = kern[r, s, ci, co]: quantised in blocks of 32 output channels
  wt = kern.permute(2, 0, 1, 3).reshape(c, 9 * cout)
  wq_ref, wsb_ref = mxfp8.quantize(wt)
  assert torch.equal(w8t, wq_ref)
  assert torch.equal(wsft, mxfp8.weight_scale_order(wsb_ref, 9, cout))
  kq = mxfp8.dequantize(wq_ref, wsb_ref).reshape(c, 3, 3, cout).permute(1, 2, 0, 3).contiguous()
  x = torch.zeros(n, h, h, c, device="cuda", requires_grad=True)
  y = O.conv2d(O.reflect_pad(x, 1), kq, 1, "valid")
  (gx,) = torch.autograd.grad(y, x, mxfp8.dequantize(q, sb))
  assert rel_err(g.t, gx) < 1.5e-2 and rel_l2(g.t, gx) < 6e-3
  x2 = torch.zeros(n, h, h, c, device="cuda", requires_grad=True)
  (gfull,) = torch.autograd.grad(O.conv2d(O.reflect_pad(x2, 1), kern, 1, "valid"), x2, dy)
  assert rel_l2(g.t, gfull) < 6e-2          # quantisation error of the E4M3 operands


def test_mxfp8_twin_of_in_bwd_apply():
  """The IN-backward-apply kernel's MXFP8 side output is the block quantisation of what it stores in bf16:
  scale bytes within one step of the reference rule, de-quantised values within E4M3 rounding of dY."""
  nv, no, O = _mods()
  from tf2_cyclegan_b200.ops import mxfp8
  C = nv.load()
  torch.manual_seed(14)
  n, h, c = 2, 32, 256
  dyt = nv.Buf(n, h, h, c); dyt.t.copy_(torch.randn(n, h, h, c, device="cuda").to(torch.bfloat16))
  raw = nv.Buf(n, h, h, c); raw.t.copy_((torch.randn(n, h, h, c, device="cuda") * 2 + 0.5).to(torch.bfloat16))
  xr = raw.t.float()
  stats = torch.stack([xr.sum((1, 2)), (xr * xr).sum((1, 2))], -1).contiguous()
  bstats = torch.randn(n, c, 2, device="cuda") * 30
  gamma = torch.rand(c, device="cuda") + 0.5
  dY = nv.Buf(n, h, h, c, halo=1)
  dY8 = nv.Buf8(n, h, h, c, halo=1)
  dsf = nv.Buf8(n, h, h, 16, halo=1)
  plan = C.Plan()
  plan.add_in_bwd_apply("bwd", dyt.view(), raw.view(), stats.data_ptr(), bstats.data_ptr(), gamma.data_ptr(),
                        1.0 / (h * h), 1e-3, dY.view(), 0, 0, n, h, h, c, dY8.view(), dsf.view())
  run(plan)
  ref = dY.interior().float()                                    # bf16 result of the same kernel
  sb_ref = mxfp8.scale_bytes(ref.reshape(n, h, h, c // 32, 32).abs().amax(-1))
  sb = dsf.interior()[..., :c // 32]
  assert (sb.int() - sb_ref.int()).abs().max().item() <= 1      # fp32 vs bf16-rounded block maximum
  back = mxfp8.dequantize(dY8.interior().contiguous(), sb.contiguous())
  s = mxfp8.scale_values(sb).repeat_interleave(32, dim=-1)
  bound = torch.maximum(ref.abs() * 2.0 ** -4, s * 2.0 ** -10) + ref.abs() * 2.0 ** -8   # E4M3 + bf16 rounding
  assert ((back - ref).abs() <= bound * 1.001).all()
  # the zero ring of the flat dgrad domain stays zero in both twins
  assert dY8.t[:, 0].abs().max() == 0 and dsf.t[:, :, 0].abs().max() == 0


@pytest.mark.parametrize("n,h,c,cout", [(2, 64, 256, 256), (3, 32, 64, 256), (1, 16, 128, 512)])
def test_conv3x3_tap_shared_boxed(n, h, c, cout):
  """conv_gemm_ts.cu: 8 x 16 pixel tiles, ONE 10 x 18 halo tile per channel block, the nine
  taps read it through row-shifted descriptors (stride between 8-row groups = 10 rows)."""
  nv, no, O = _mods()
  C = nv.load()
  torch.manual_seed(10)
  x = bf(torch.randn(n, h, h, c, device="cuda"))
  k = bf(torch.randn(3, 3, c, cout, device="cuda") * 0.05)
  xin = nv.Buf(n, h, h, c, halo=1)
  fill_reflect(xin, x)
  w2 = no.pack_conv_weight(k).to(torch.bfloat16)
  out = nv.Buf(n, h, h, cout)
  stats = torch.zeros(n, cout, 2, device="cuda")
  plan = C.Plan()
  no.add_conv(plan, "conv", [xin.tmap_padded(8, 16)], no.taps_conv(3, c), c, w2, out.view(), h, h,
              n, cout, stats=stats.data_ptr(), tile=(8, 16), halo_map=lambda w_, h_: xin.tmap_padded(w_, h_))
  run(plan)
  ref = O.conv2d(O.reflect_pad(x, 1), k, 1, "valid")
  assert rel_err(out.t, ref) < 1e-2
  got = out.t.float()
  assert torch.allclose(stats[..., 0], got.sum((1, 2)), rtol=1e-3, atol=1e-2 * h)
  assert torch.allclose(stats[..., 1], (got * got).sum((1, 2)), rtol=1e-3, atol=1e-2 * h)


def test_conv4x4_same_tap_shared_zero_fill():
  """TF-SAME 4x4 stride-1 conv (PatchGAN conv3 shape): the halo box starts at negative
  coordinates, TMA zero-fills outside the image."""
  nv, no, O = _mods()
  C = nv.load()
  torch.manual_seed(11)
  n, h, c, cout = 2, 32, 64, 512
  x = bf(torch.randn(n, h, h, c, device="cuda"))
  k = bf(torch.randn(4, 4, c, cout, device="cuda") * 0.05)
  xin = nv.Buf(n, h, h, c)
  xin.t.copy_(x.to(torch.bfloat16))
  w2 = no.pack_conv_weight(k).to(torch.bfloat16)
  out = nv.Buf(n, h, h, cout)
  plan = C.Plan()
  no.add_conv(plan, "conv", [xin.tmap_interior(8, 16)], no.taps_conv(4, c, origin=-1), c, w2, out.view(),
              h, h, n, cout, tile=(8, 16), halo_map=lambda w_, h_: xin.tmap_interior(w_, h_))
  run(plan)
  ref = O.conv2d(x, k, 1, "same")
  assert rel_err(out.t, ref) < 1e-2


def test_dgrad_flat_tap_shared():
  """flat-domain dgrad with one 272-row halo tile (two TMA boxes) shared by the nine taps,
  MN-major B = forward weights."""
  nv, no, O = _mods()
  C = nv.load()
  torch.manual_seed(12)
  n, h, c, cout = 2, 32, 256, 128
  x = bf(torch.randn(n, h, h, c, device="cuda")).requires_grad_(True)
  kern = bf(torch.randn(3, 3, c, cout, device="cuda") * 0.05)
  dy = bf(torch.randn(n, h, h, cout, device="cuda"))
  y = O.conv2d(O.reflect_pad(x, 1), kern, 1, "valid")
  (gx,) = torch.autograd.grad(y, x, dy)
  dyb = nv.Buf(n, h, h, cout, halo=1)
  dyb.interior().copy_(dy.to(torch.bfloat16))
  w2 = no.pack_conv_weight(kern).to(torch.bfloat16)
  P = h + 2
  rows = n * P * P
  dxpad = nv.Buf(n, h, h, c, halo=1)
  taps = [(0, 0, (1 - r) * P + (1 - s), 0, (r * 3 + s) * c) for r in range(3) for s in range(3)]
  plan = C.Plan()
  no.add_conv_flat(plan, "dgrad", dyb.tmap_flat(), taps, cout, w2, dxpad.ptr(padded=True), c, rows, c,
                   b_mn=True, halo_map=lambda rb: dyb.tmap_flat(rows_box=rb))
  g = nv.Buf(n, h, h, c)
  plan.add_in_bwd_reduce("fold", dxpad.view(), 1, nv.NULL_VIEW, nv.NULL_VIEW, 0, 0, 0, 1.0, 1e-3, 0,
                         g.view(), 0, n, h, h, c)
  run(plan)
  assert rel_err(g.t, gx) < 1.5e-2
