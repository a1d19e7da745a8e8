"""Stand-in baseline — NOT the reference.

The reference (bryanlimy/tf2-cyclegan, TF 2.4.1 + tensorflow-addons + TFDS, CUDA 11.0) cannot be
installed on this Python-3.12 offline image, so ``bench.py --impl reference``
prints ``unavailable``. This file is the strongest *library-kernel* implementation of the same train
step we could write with stock PyTorch, i.e. the same division of labour as the reference's
MirroredStrategy(NCCL) + cuDNN build, tuned the way a careful PyTorch user would tune it:

  * NCHW tensors in ``channels_last`` memory format, bf16 compute (weights kept as fp32 masters with
    bf16 compute copies), cuDNN convolutions with ``cudnn.benchmark`` autotuning;
  * ATen's fused instance-norm kernels (``F.instance_norm``), ``F.pad(mode="reflect")``;
  * the reference's exact schedule and gradient routing (main.py:207-262): 6 generator + 6
    discriminator forwards under one autograd graph, four ``autograd.grad`` calls (one per
    optimizer's var_list), un-detached fakes, losses pre-divided by the GLOBAL batch;
  * gradients land in ONE flat fp32 buffer per network -> one packed NCCL ``all_reduce(SUM)`` per
    network (the zero-copy form of ``NcclAllReduce(num_packs=1)``), then a multi-tensor
    (``torch._foreach``) Keras-formula Adam (lr 2e-4, beta1 0.5, beta2 0.9, eps 1e-7 outside the
    bias correction) and a foreach refresh of the bf16 copies;
  * the whole step captured in CUDA graphs (one graph on 1 GPU; forward + backward and the optimizers
    as two graphs around the four NCCL calls on several GPUs).

It shares no code with the product engine (``tf2_cyclegan_b200``): plain ``torch.nn.functional``
calls only. Architecture per cyclegan/model.py:129-213 (TF-SAME asymmetric padding, transposed-conv
crop, IN eps 1e-3, gamma ~ N(0, 0.02), LeakyReLU 0.2, PatchGAN without norm on the first layer).
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch
import torch.distributed as dist
import torch.nn.functional as F

IN_EPS = 1e-3


# ----------------------------------------------------------------------------- parameters
def _normal(shape, gen, std=0.02):
  return torch.randn(shape, generator=gen) * std


def generator_params(gen: torch.Generator, filters=64, n_res=9) -> List[Tuple[str, torch.Tensor]]:
  p: List[Tuple[str, torch.Tensor]] = []

  def conv(name, cout, cin, k, norm=True, transpose=False):
    shape = (cin, cout, k, k) if transpose else (cout, cin, k, k)
    p.append((f"{name}.w", _normal(shape, gen)))
    if norm:
      p.append((f"{name}.gamma", _normal((cout,), gen)))
      p.append((f"{name}.beta", torch.zeros(cout)))

  conv("stem", filters, 3, 7)
  c = filters
  for i in range(2):
    conv(f"down{i}", 2 * c, c, 3)
    c *= 2
  for i in range(n_res):
    conv(f"res{i}a", c, c, 3)
    conv(f"res{i}b", c, c, 3)
  for i in range(2):
    conv(f"up{i}", c // 2, c, 3, transpose=True)
    c //= 2
  bound = math.sqrt(6.0 / (7 * 7 * c + 7 * 7 * 3))       # glorot_uniform (Keras default) on the output conv
  p.append(("final.w", (torch.rand((3, c, 7, 7), generator=gen) * 2 - 1) * bound))
  p.append(("final.b", torch.zeros(3)))
  return p


def discriminator_params(gen: torch.Generator, filters=64) -> List[Tuple[str, torch.Tensor]]:
  p: List[Tuple[str, torch.Tensor]] = [("conv0.w", _normal((filters, 3, 4, 4), gen)), ("conv0.b", torch.zeros(filters))]
  c = filters
  for i in range(3):
    p.append((f"down{i}.w", _normal((2 * c, c, 4, 4), gen)))
    p.append((f"down{i}.gamma", _normal((2 * c,), gen)))
    p.append((f"down{i}.beta", torch.zeros(2 * c)))
    c *= 2
  p.append(("out.w", _normal((1, c, 4, 4), gen)))
  p.append(("out.b", torch.zeros(1)))
  return p


# ------------------------------------------------------------------------------- forward
def _same_pad(x, k, s):
  """TF 'SAME' zero padding (asymmetric: the extra pixel goes after)."""
  h, w = x.shape[-2:]
  ph = max((math.ceil(h / s) - 1) * s + k - h, 0)
  pw = max((math.ceil(w / s) - 1) * s + k - w, 0)
  return F.pad(x, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))


def _in(x, gamma, beta):
  return F.instance_norm(x, weight=gamma, bias=beta, eps=IN_EPS)


def generator_forward(P: Dict[str, torch.Tensor], x: torch.Tensor, n_res: int) -> torch.Tensor:
  h = F.conv2d(F.pad(x, (3, 3, 3, 3), mode="reflect"), P["stem.w"])
  h = F.relu(_in(h, P["stem.gamma"], P["stem.beta"]))
  for i in range(2):
    h = F.conv2d(_same_pad(h, 3, 2), P[f"down{i}.w"], stride=2)
    h = F.relu(_in(h, P[f"down{i}.gamma"], P[f"down{i}.beta"]))
  for i in range(n_res):
    r = F.conv2d(F.pad(h, (1, 1, 1, 1), mode="reflect"), P[f"res{i}a.w"])
    r = F.relu(_in(r, P[f"res{i}a.gamma"], P[f"res{i}a.beta"]))
    r = F.conv2d(F.pad(r, (1, 1, 1, 1), mode="reflect"), P[f"res{i}b.w"])
    h = h + _in(r, P[f"res{i}b.gamma"], P[f"res{i}b.beta"])
  for i in range(2):
    # Conv2DTranspose(k3, s2, "same") = full transposed conv cropped by (0, 1)
    h = F.conv_transpose2d(h, P[f"up{i}.w"], stride=2)[..., :-1, :-1]
    h = F.relu(_in(h, P[f"up{i}.gamma"], P[f"up{i}.beta"]))
  h = F.conv2d(F.pad(h, (3, 3, 3, 3), mode="reflect"), P["final.w"], P["final.b"])
  return torch.tanh(h)


def discriminator_forward(P: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
  h = F.leaky_relu(F.conv2d(_same_pad(x, 4, 2), P["conv0.w"], P["conv0.b"], stride=2), 0.2)
  for i in range(3):
    s = 2 if i < 2 else 1
    h = F.conv2d(_same_pad(h, 4, s), P[f"down{i}.w"], stride=s)
    h = F.leaky_relu(_in(h, P[f"down{i}.gamma"], P[f"down{i}.beta"]), 0.2)
  return F.conv2d(_same_pad(h, 4, 1), P["out.w"], P["out.b"])


# ------------------------------------------------------------------------------- trainer
class FlatNet:
  """fp32 master / Adam moments / fp32 gradients of one network as flat buffers; bf16 compute
  copies are the autograd leaves (channels_last for 4-D kernels)."""

  def __init__(self, named: List[Tuple[str, torch.Tensor]], device):
    self.names = [n for n, _ in named]
    self.shapes = [t.shape for _, t in named]
    sizes = [t.numel() for _, t in named]
    self.n = sum(sizes)
    self.master = torch.cat([t.reshape(-1).float() for _, t in named]).to(device)
    self.m = torch.zeros_like(self.master)
    self.v = torch.zeros_like(self.master)
    self.grad = torch.zeros_like(self.master)
    offs = [0]
    for s in sizes:
      offs.append(offs[-1] + s)
    self.master_views = [self.master[a:b].view(sh) for a, b, sh in zip(offs, offs[1:], self.shapes)]
    self.grad_views = [self.grad[a:b].view(sh) for a, b, sh in zip(offs, offs[1:], self.shapes)]
    self.leaves = []
    for mv in self.master_views:
      t = mv.to(torch.bfloat16)
      if t.dim() == 4:
        t = t.contiguous(memory_format=torch.channels_last)
      self.leaves.append(t.requires_grad_(True))
    self.P = dict(zip(self.names, self.leaves))
    self.step = torch.zeros((), dtype=torch.int64, device=device)

  def adam(self, lr=2e-4, b1=0.5, b2=0.9, eps=1e-7):
    """Keras Adam on the flat buffers (one multi-tensor pass), then refresh the bf16 copies. The step
    counter and the bias-corrected rate live on the device so the update is CUDA-graph replayable."""
    self.step += 1
    t = self.step.float()
    lr_t = lr * torch.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
    g = self.grad
    self.m.mul_(b1).add_(g, alpha=1.0 - b1)
    self.v.mul_(b2).addcmul_(g, g, value=1.0 - b2)
    self.master.addcdiv_(self.m * lr_t, self.v.sqrt().add_(eps), value=-1.0)
    with torch.no_grad():
      torch._foreach_copy_([l.data for l in self.leaves], self.master_views)


class StandinCycleGAN:

  def __init__(self, device, local_batch: int, global_batch: int, image_size: int = 256, n_res: int = 9,
               seed: int = 1234, world: int = 1):
    self.dev, self.b, self.gbs, self.S, self.n_res, self.world = device, local_batch, global_batch, image_size, n_res, world
    gen = torch.Generator().manual_seed(seed)
    self.nets = {"G": FlatNet(generator_params(gen, n_res=n_res), device),
                 "F": FlatNet(generator_params(gen, n_res=n_res), device),
                 "X": FlatNet(discriminator_params(gen), device),
                 "Y": FlatNet(discriminator_params(gen), device)}
    self.x = torch.zeros(local_batch, 3, image_size, image_size, device=device, dtype=torch.bfloat16).contiguous(
        memory_format=torch.channels_last)
    self.y = torch.zeros_like(self.x)
    self.losses = torch.zeros(10, device=device)
    self.graph = None

  # losses: per-sample mean over non-batch axes, SUM over the local batch / GLOBAL batch (main.py:172-195)
  def _mae(self, a, b):
    return (a.float() - b.float()).abs().mean(dim=(1, 2, 3)).sum() / self.gbs

  def _mse(self, target, d):
    return ((target - d.float()) ** 2).mean(dim=(1, 2, 3)).sum() / self.gbs

  def _fwd_bwd(self):
    """Forward + the four backward passes; gradients land in the flat fp32 buffers."""
    N = self.nets
    G = lambda t: generator_forward(N["G"].P, t, self.n_res)
    Fn = lambda t: generator_forward(N["F"].P, t, self.n_res)
    X = lambda t: discriminator_forward(N["X"].P, t)
    Y = lambda t: discriminator_forward(N["Y"].P, t)
    x, y = self.x, self.y
    fake_y, fake_x = G(x), Fn(y)
    dx_fake, dy_fake = X(fake_x), Y(fake_y)           # forwarded once, used by both losses
    G_loss, F_loss = self._mse(1.0, dy_fake), self._mse(1.0, dx_fake)
    G_cycle, F_cycle = 10.0 * self._mae(y, G(fake_x)), 10.0 * self._mae(x, Fn(fake_y))
    G_id, F_id = 5.0 * self._mae(y, G(y)), 5.0 * self._mae(x, Fn(x))
    G_total, F_total = G_loss + G_cycle + G_id, F_loss + F_cycle + F_id
    X_loss = 0.5 * (self._mse(1.0, X(x)) + self._mse(0.0, dx_fake))
    Y_loss = 0.5 * (self._mse(1.0, Y(y)) + self._mse(0.0, dy_fake))
    for name, loss in (("G", G_total), ("F", F_total), ("X", X_loss), ("Y", Y_loss)):
      net = N[name]
      grads = torch.autograd.grad(loss, net.leaves, retain_graph=True)
      torch._foreach_copy_(net.grad_views, list(grads))          # bf16 -> flat fp32 gradient buffer
    self.losses.copy_(torch.stack([G_loss, G_cycle, G_id, G_total, F_loss, F_cycle, F_id, F_total, X_loss, Y_loss]).detach())

  def _reduce(self):
    if self.world > 1:
      for name in ("G", "F", "X", "Y"):
        dist.all_reduce(self.nets[name].grad)                      # one packed SUM per network

  def _optimize(self):
    for name in ("G", "F", "X", "Y"):
      self.nets[name].adam()

  def _step_body(self):
    self._fwd_bwd()
    self._reduce()
    self._optimize()

  def capture(self) -> bool:
    """Warm up on a side stream (cuDNN autotune, lazy init), then capture the step in CUDA graphs:
    one graph on a single GPU; on several GPUs two graphs (forward + backward, optimizers) with the
    four packed NCCL all-reduces issued between them (NCCL calls are kept out of the capture)."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
      for _ in range(3):
        self._step_body()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    try:
      if self.world == 1:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
          self._step_body()
        self.graph = (g,)
      else:
        g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1):
          self._fwd_bwd()
        with torch.cuda.graph(g2, pool=g1.pool()):
          self._optimize()
        self.graph = (g1, g2)
      return True
    except Exception as e:  # pragma: no cover - depends on the torch build
      print(f"[standin] CUDA-graph capture failed ({type(e).__name__}: {e}); running eagerly")
      self.graph = None
      torch.cuda.synchronize()
      return False

  def train_step(self, x_nhwc: torch.Tensor, y_nhwc: torch.Tensor) -> torch.Tensor:
    """x, y: [b, H, W, 3] fp32 on the device. Returns the 10 loss scalars (device tensor)."""
    # NHWC fp32 -> logical NCHW bf16 in channels_last memory: a cast, no transpose
    self.x.copy_(x_nhwc.permute(0, 3, 1, 2))
    self.y.copy_(y_nhwc.permute(0, 3, 1, 2))
    if self.graph is None:
      self._step_body()
    elif len(self.graph) == 1:
      self.graph[0].replay()
    else:
      self.graph[0].replay()
      self._reduce()
      self.graph[1].replay()
    return self.losses
