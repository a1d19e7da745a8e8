"""Native (sm_100a) backend: the whole CycleGAN train step as one pre-compiled
Plan of hand-written kernels — no autograd, no cuDNN/cuBLAS, no per-step Python
on the hot path (CUDA-graph replay).

Schedule (exploits the reference's gradient routing):
  * each generator holds three image slots: G: [G(x) | G(y) | G(fake_x)],
    F: [F(y) | F(x) | F(fake_y)]. Slots 0-1 are ONE batched forward, slot 2 a second;
    the backward (dgrad + InstanceNorm backward + wgrad) is ONE batched pass over 3b
    images, so the three calls' weight gradients accumulate inside one split-K GEMM;
  * discriminators see [D(fake) | D(real)] as one batch; the fakes are forwarded once
    (the reference forwards them twice, main.py:213-214,241-242);
  * the adversarial term reaches a generator through a dgrad-only pass of the
    discriminator; no dgrad is ever computed into cycle inputs or past a stem;
  * all four updates use pre-update weights; G's and F's optimizer steps are
    issued right after their own backward.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Optional

import torch

from ..config import ENGINE
from ..models import oracle_net
from ..models.spec import NetSpec
from ..ops import native as nv
from ..ops.native import NULL_VIEW, Buf
from ..parallel.strategy import Strategy
from .native_net import NetExec, NetState

NETS = ("G", "F", "X", "Y")
SLOT = {"G_loss": 0, "G_cycle": 1, "G_identity": 2, "F_loss": 3, "F_cycle": 4, "F_identity": 5,
        "X_loss": 6, "Y_loss": 7, "err_x_cycle": 8, "err_y_cycle": 9, "err_x_same": 10,
        "err_y_same": 11}
NUM_SLOTS = 16
COMM_LANE = 4          # csrc/plan.cu: the high-priority stream that carries every cross-GPU launch
SCALAR_NET = "X"       # the network whose last fused launch also carries the loss-scalar sums
ADAM = dict(lr=2e-4, b1=0.5, b2=0.9, eps=1e-7)   # main.py:134-145 (+ Keras default epsilon)


class StepPlan:
  """Everything that depends on the local batch size b."""

  def __init__(self, be: "NativeBackend", b: int):
    C = nv.load()
    self.b = b
    dev = be.device
    S = be.image_hw
    IH, IW = S
    self.x_in = torch.zeros(b, IH, IW, 3, device=dev)
    self.y_in = torch.zeros(b, IH, IW, 3, device=dev)
    self.slots = be.slots          # loss / metric accumulators (peer-readable on the fused path)
    self.ex: Dict[str, NetExec] = {
        "G": NetExec(be.specs["G"], be.state["G"], 3 * b, S, dev, "G", fp8=be.fp8, pool=be.buffer_pool),
        "F": NetExec(be.specs["F"], be.state["F"], 3 * b, S, dev, "F", fp8=be.fp8, pool=be.buffer_pool),
        "X": NetExec(be.specs["X"], be.state["X"], 2 * b, S, dev, "X", pool=be.buffer_pool),
        "Y": NetExec(be.specs["Y"], be.state["Y"], 2 * b, S, dev, "Y", pool=be.buffer_pool),
    }
    for e in self.ex.values():
      e.enable_backward()
    G, F, X, Y = (self.ex[n] for n in NETS)
    hd, wd = X.out_hw[-1]
    self.gd_adv = {"X": Buf(b, hd, wd, 8, device=dev), "Y": Buf(b, hd, wd, 8, device=dev)}
    gbs = be.global_batch_size
    plan = C.Plan()
    plan.set_nvtx(ENGINE.nvtx)
    self.plan = plan
    self.marks: Dict[str, int] = {}
    sl = lambda k: self.slots[SLOT[k]:].data_ptr()

    # ---- P0: zero accumulators
    plan.add_memset("zero.slots", self.slots.data_ptr(), NUM_SLOTS * 4)
    for n in NETS:
      e = self.ex[n]
      plan.add_memset(f"zero.{n}.stats", e.stat_arena.data_ptr(), e.stat_arena.numel() * 4)
      plan.add_memset(f"zero.{n}.bstats", e.bstat_arena.data_ptr(), e.bstat_arena.numel() * 4)
    if be.fp8:   # FP8 weight copies + scales for this step (from the previous step's statistics)
      for n in ("G", "F"):
        be.state[n].emit_fp8_refresh(plan, n)
    # ---- P1: inputs -> 8-channel padded bf16 layouts
    xi, yi = self.x_in.data_ptr(), self.y_in.data_ptr()
    plan.add_img_prep("prep.x->G0,X1", xi, NULL_VIEW, G.inp[0].view(0), 3, X.inp[0].view(b), b, IH, IW)
    plan.add_img_prep("prep.x->F1", xi, NULL_VIEW, F.inp[0].view(b), 3, NULL_VIEW, b, IH, IW)
    plan.add_img_prep("prep.y->F0,Y1", yi, NULL_VIEW, F.inp[0].view(0), 3, Y.inp[0].view(b), b, IH, IW)
    plan.add_img_prep("prep.y->G1", yi, NULL_VIEW, G.inp[0].view(b), 3, NULL_VIEW, b, IH, IW)
    # Two lanes: lane 0 = {G, Y}, lane 1 = {F, X}. Inside a fork/join region the lanes run
    # on two streams, so one network's tail wave overlaps the other's kernels.
    two = be.two_lanes

    def fork(tag):
      if two:
        plan.add_fork(f"sync.fork.{tag}")

    def join(tag):
      plan.set_lane(0)
      if two:
        plan.add_join(f"sync.join.{tag}")

    def lane(k):
      plan.set_lane(k if two else 0)

    # ---- P2: fake / identity generators (one batched pass each)
    fork("fwd")
    lane(0); G.forward(plan, 0, 2 * b)
    lane(1); F.forward(plan, 0, 2 * b)
    join("fwd")
    # ---- P3: fakes -> inputs of the cycle pass and of the discriminators
    plan.add_img_prep("prep.fake_y->F2,Y0", 0, G.inp[-1].view(0), F.inp[0].view(2 * b), 3,
                      Y.inp[0].view(0), b, IH, IW)
    plan.add_img_prep("prep.fake_x->G2,X0", 0, F.inp[-1].view(0), G.inp[0].view(2 * b), 3,
                      X.inp[0].view(0), b, IH, IW)
    # ---- P4: cycle pass
    fork("cycle")
    lane(0); G.forward(plan, 2 * b, b)
    lane(1); F.forward(plan, 2 * b, b)
    join("cycle")
    self.marks["cycle_end"] = plan.size()
    # ---- P5: discriminators on [fake | real]
    fork("dis")
    lane(0); Y.forward(plan, 0, 2 * b)
    lane(1); X.forward(plan, 0, 2 * b)
    join("dis")
    # ---- P6: losses (+ their gradients)
    npx = IH * IW * 3
    nd = hd * wd
    la = be.lambda_adv
    plan.add_mse_loss("loss.G_adv", Y.inp[-1].view(0), 1.0, la / (nd * gbs), 2.0 * la / (nd * gbs),
                      sl("G_loss"), self.gd_adv["Y"].view(0), b, hd, wd)
    plan.add_mse_loss("loss.F_adv", X.inp[-1].view(0), 1.0, la / (nd * gbs), 2.0 * la / (nd * gbs),
                      sl("F_loss"), self.gd_adv["X"].view(0), b, hd, wd)
    lc, li = be.lambda_cycle, be.lambda_identity
    plan.add_l1_loss("loss.G_cycle", G.inp[-1].view(2 * b), yi, lc / (npx * gbs), lc / (npx * gbs),
                     sl("G_cycle"), sl("err_y_cycle"), 1.0 / (npx * gbs), G.gout_final.view(2 * b), b, IH, IW)
    plan.add_l1_loss("loss.G_identity", G.inp[-1].view(b), yi, li / (npx * gbs), li / (npx * gbs),
                     sl("G_identity"), sl("err_y_same"), 1.0 / (npx * gbs), G.gout_final.view(b), b, IH, IW)
    plan.add_l1_loss("loss.F_cycle", F.inp[-1].view(2 * b), xi, lc / (npx * gbs), lc / (npx * gbs),
                     sl("F_cycle"), sl("err_x_cycle"), 1.0 / (npx * gbs), F.gout_final.view(2 * b), b, IH, IW)
    plan.add_l1_loss("loss.F_identity", F.inp[-1].view(b), xi, li / (npx * gbs), li / (npx * gbs),
                     sl("F_identity"), sl("err_x_same"), 1.0 / (npx * gbs), F.gout_final.view(b), b, IH, IW)
    for n in ("X", "Y"):
      D = self.ex[n]
      plan.add_mse_loss(f"loss.{n}_real", D.inp[-1].view(b), 1.0, 0.5 / (nd * gbs), 1.0 / (nd * gbs),
                        sl(f"{n}_loss"), D.gout_final.view(b), b, hd, wd)
      plan.add_mse_loss(f"loss.{n}_fake", D.inp[-1].view(0), 0.0, 0.5 / (nd * gbs), 1.0 / (nd * gbs),
                        sl(f"{n}_loss"), D.gout_final.view(0), b, hd, wd)
    self.marks["test_end"] = plan.size()

    # ---- P7-P9: backward + optimizer. lane 0: Y adversarial dgrad -> G backward -> G update ->
    # Y backward -> Y update; lane 1: the same for X / F. (sequential when lanes are off,
    # which is what the NCCL baseline path needs: one all-reduce between bwd and optimizer)
    def adv(D, gen):
      e = self.ex[D]
      e.backward(plan, 0, b, wgrad=False, input_grad_target=(self.ex[gen].gout_final, 0),
                 gout_view=self.gd_adv[D].view(0), gout_is_dy=True)
      plan.add_memset(f"zero.{D}.bstats2", e.bstat_arena.data_ptr(), e.bstat_arena.numel() * 4)

    # Fused multi-GPU path: the optimizer (cross-GPU reduce + Adam + weight broadcast) is issued in
    # BUCKETS while the backward pass is still running. Every cross-GPU launch goes to ONE dedicated
    # lane (COMM_LANE) in ONE global order that is identical on every rank — including ranks with an
    # empty share of the batch, which replay just this sequence — so two ranks can never wait on each
    # other's not-yet-launched kernel in opposite orders. Only the last launch of each network (its
    # first layers, the one-shot vector tail and, for the last network, the loss scalars) is exposed.
    bucketed = two and be.fused
    comm_requests = []   # (network, bucket index, [event ops])

    def full(n, lanes=None):
      e = self.ex[n]
      hook = None
      if bucketed:
        bs = be.buckets[n]
        first = {j: k for k, (j, _, _) in enumerate(bs)}
        main_lane = lanes[0] if lanes is not None else (0 if n in ("G", "Y") else 1)

        def hook(j):
          k = first.get(j)
          if k is None:
            return
          evs = [plan.add_event_record(f"sync.rec.{n}.bucket{k}.main", main_lane)]
          if lanes is not None:
            evs.append(plan.add_event_record(f"sync.rec.{n}.bucket{k}.wgrad", lanes[1]))
          comm_requests.append((n, k, evs))

      if n in ("G", "F"):
        e.backward(plan, 0, 3 * b, wgrad=True, lanes=lanes, join_wgrad=not bucketed, on_layer_done=hook)
      else:
        e.backward(plan, 0, 2 * b, wgrad=True, gout_is_dy=True, lanes=lanes, join_wgrad=not bucketed,
                   on_layer_done=hook)
      self.marks[f"{n}_bwd_end"] = plan.size()
      if not bucketed:
        be.emit_optimizer(plan, n)
      self.marks[f"{n}_opt_end"] = plan.size()

    if two:
      fork("bwd")
      wl = be.wgrad_lanes
      lane(0); adv("Y", "G"); full("G", (0, 2) if wl else None); full("Y", (0, 2) if wl else None)
      lane(1); adv("X", "F"); full("F", (1, 3) if wl else None); full("X", (1, 3) if wl else None)
      if bucketed:
        ready = {(n, k): evs for n, k, evs in comm_requests}
        assert len(ready) == len(be.comm_order()), (sorted(ready), be.comm_order())
        plan.set_lane(COMM_LANE)
        for n, k in be.comm_order():
          for ev in ready[(n, k)]:
            plan.add_event_wait(f"sync.wait.{n}.bucket{k}", COMM_LANE, ev)
          be.emit_optimizer(plan, n, bucket=k)
      join("bwd")
    else:
      adv("Y", "G"); adv("X", "F")
      for n in NETS:
        full(n)
    self.marks["end"] = plan.size()
    self.names = plan.names()
    self.kernels_per_step = sum(1 for nm in self.names
                                if not nm.startswith(("zero.", "sync.")) and ".zero" not in nm)
    self.graph: Optional[torch.cuda.CUDAGraph] = None


class NativeBackend:
  name = "native"

  def __init__(self, specs: Dict[str, NetSpec], global_batch_size: int, strategy: Strategy,
               seed: int = 1234, image_size: int = 256, local_batch: int = 1, dtype: str = "bf16",
               fused_allreduce: str = "auto", use_graph: Optional[bool] = None):
    if not nv.available():
      raise RuntimeError("the native backend needs an sm_100 GPU (B200)")
    if dtype not in ("bf16", "fp8", "mxfp8"):
      raise NotImplementedError(f"native backend: dtype {dtype!r} is not available (bf16 | fp8 | mxfp8)")
    # residual-block forward convs in E4M3: "tensor" = one delayed scale per tensor (--dtype fp8),
    # "mx" = OCP MXFP8, one E8M0 scale per 32 channels of either operand (--dtype mxfp8); "" = bf16
    self.fp8 = {"fp8": "tensor", "mxfp8": "mx"}.get(dtype, "")
    hw = (image_size, image_size) if isinstance(image_size, int) else tuple(int(v) for v in image_size)
    if len(hw) != 2 or min(hw) < 32 or hw[0] % 8 or hw[1] % 8:
      # three stride-2 stages in the discriminator (and two in the generators) must see even inputs;
      # anything else the model admits (cyclegan/model.py:129: input_shape is free) runs natively:
      # tiles are 128 pixels wide x high with a masked partial last tile row
      raise ValueError("native backend: image height and width must be multiples of 8 and >= 32; "
                       "use --backend oracle for other shapes")
    nv.load()
    self.specs = specs
    self.strategy = strategy
    self.device = strategy.device
    self.global_batch_size = global_batch_size
    self.image_size = image_size
    self.image_hw = hw
    self.lambda_cycle = 10.0
    self.lambda_identity = 5.0
    self.lambda_adv = 1.0     # weight of the generators' LSGAN term: 1 in the reference; a knob for tests
    self.use_graph = ENGINE.use_graph if use_graph is None else use_graph
    self.world = strategy.world_size
    self.fused = (fused_allreduce in ("auto", "on")) and self.world > 1
    self.two_lanes = ENGINE.lanes >= 2
    self.wgrad_lanes = ENGINE.lanes >= 4
    gen = torch.Generator().manual_seed(seed)
    self.state: Dict[str, NetState] = OrderedDict()
    self.symm = None
    if self.world > 1 and not (fused_allreduce in ("auto", "on")):
      self.two_lanes = False   # NCCL baseline path runs plan segments sequentially
    if self.fused:
      from ..parallel.fused_allreduce import SymmetricBuffers
      try:
        self.symm = SymmetricBuffers(strategy)
      except Exception as e:  # pragma: no cover - depends on the box
        if fused_allreduce == "on":
          raise
        print(f"[native] fused all-reduce unavailable ({type(e).__name__}: {e}); using NCCL path")
        self.fused = False
        self.two_lanes = False
    for n in NETS:
      logical = oracle_net.init_params(specs[n], generator=gen)
      gbuf = wbuf = None
      if self.fused:
        from .native_params import FlatLayout
        numel = FlatLayout(specs[n]).numel
        gbuf, wbuf = self.symm.alloc(n, numel)
      self.state[n] = NetState(specs[n], logical, self.device, grad_buffer=gbuf, wb_buffer=wbuf)
    # gradient all-reduce buckets (fused path): plan order = the order their gradients become final
    self.buckets: Dict[str, List[tuple]] = {}
    self.bucket_index: Dict[str, List[int]] = {}
    if self.fused:
      for n in NETS:
        bs = self.plan_buckets(n)
        asc = sorted((lo, hi) for _, lo, hi in bs)
        self.symm.set_buckets(n, asc, vec_lo=self.state[n].layout.vec_lo)
        self.buckets[n] = bs
        self.bucket_index[n] = [asc.index((lo, hi)) for _, lo, hi in bs]
      # push-mode staging buffer, sized once for the largest bucket (symmetric allocation: collective)
      self.symm.stage_ptrs(max(hi - lo for n in NETS for _, lo, hi in self.buckets[n]))
    # loss / metric slots: on the fused path they live in peer-readable memory and the last fused
    # launch of a step leaves their sums over ranks in symm.scalars_sum (no NCCL call, no host sync)
    self.slots = self.symm.scalars if self.fused else torch.zeros(NUM_SLOTS, device=self.device)
    self.buffer_pool: dict = {}     # activation / gradient arenas shared by the plans of all batch sizes
    self._opt_only_plan = None
    self._scalar_plan = None
    self._plans: Dict[int, StepPlan] = {}
    self._launches = 0
    torch.cuda.synchronize(self.device)
    strategy.barrier()

  # ------------------------------------------------------------ optimizer
  def plan_buckets(self, n: str) -> List[tuple]:
    """All-reduce buckets of network n: [(first layer index, lo, hi)] in the order their gradients
    become final (the backward pass visits layers last to first), about equal in bytes. A bucket
    is reduced as soon as its first (lowest) layer has been processed."""
    st = self.state[n]
    layers = st.layout.spec.layers
    end = st.layout.vec_lo                      # kernels only; the vector tail is one-shot
    offs = [st.layout.seg(l, "kernel").offset for l in layers] + [end]
    want = 4 if n in ("G", "F") else 2
    target = end / max(1, want)
    out, hi = [], end
    for j in range(len(layers) - 1, -1, -1):
      if j == 0 or (hi - offs[j] >= target and len(out) < want - 1):
        out.append((j, offs[j] if j else 0, hi))
        hi = offs[j]
    return out

  def comm_order(self) -> List[tuple]:
    """Global order of the fused cross-GPU launches of one step: [(network, bucket index)], the
    generators' buckets interleaved (they become ready alternately on the two main lanes), then
    the discriminators'. Identical on every rank by construction."""
    out = []
    for pair in (("G", "F"), ("Y", "X")):
      for k in range(max(len(self.buckets[n]) for n in pair)):
        out += [(n, k) for n in pair if k < len(self.buckets[n])]
    return out

  def emit_optimizer(self, plan, n: str, bucket: Optional[int] = None):
    """Optimizer ops of network n. Fused path: ``bucket`` selects ONE all-reduce bucket (index into
    plan_buckets(n)); the network's last bucket also updates the one-shot vector tail, carries the
    loss-scalar sums (SCALAR_NET) and is followed by the per-step tail (step counter, derived weights)."""
    st = self.state[n]
    if self.fused:
      ks = range(len(self.buckets[n])) if bucket is None else [bucket]
      for k in ks:
        finish = k == len(self.buckets[n]) - 1
        _, lo, hi = self.buckets[n][k]
        plan.add_mask_grad(f"opt.{n}.mask[{lo}:{hi}]", st.g.data_ptr() + 4 * lo, st.mask.data_ptr() + lo, hi - lo)
        self.symm.emit_allreduce_adam(plan, n, st, bucket=self.bucket_index[n][k], vectors=finish,
                                      scalars=finish and n == SCALAR_NET, exposed=finish, **ADAM)
        if finish:
          plan.add_step_increment(f"opt.{n}.step", st.step.data_ptr())
          st.refresh_derived_ops(plan, f"opt.{n}")
      return
    plan.add_mask_grad(f"opt.{n}.mask", st.g.data_ptr(), st.mask.data_ptr(), st.n)
    plan.add_adam(f"opt.{n}.adam", st.w.data_ptr(), st.m.data_ptr(), st.v.data_ptr(), st.g.data_ptr(),
                  st.wb.data_ptr(), st.step.data_ptr(), ADAM["lr"], ADAM["b1"], ADAM["b2"], ADAM["eps"],
                  1.0, st.n)
    plan.add_memset(f"zero.{n}.grads", st.g.data_ptr(), st.n * 4)
    plan.add_step_increment(f"opt.{n}.step", st.step.data_ptr())
    st.refresh_derived_ops(plan, f"opt.{n}")

  # ---------------------------------------------------------------- plans
  def plan_for(self, b: int) -> StepPlan:
    if b not in self._plans:
      self._plans[b] = StepPlan(self, b)
    return self._plans[b]

  def kernel_launches(self) -> int:
    return self._launches

  def _stream(self) -> int:
    return torch.cuda.current_stream(self.device).cuda_stream

  @property
  def reduces_internally(self) -> bool:
    """True when step results are already summed over replicas (fused path: loss-scalar side
    channel of the last fused launch) — callers must not reduce them again."""
    return self.fused

  def reduce_results(self, d: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """SUM step results over replicas (main.py:264-267): a no-op on the fused path, one packed
    NCCL all-reduce on the baseline path."""
    return d if self.fused else self.strategy.reduce_dict(d)

  def collective(self) -> str:
    """'fused-<push|pull>-<nvls|p2p>' | 'nccl' | 'none' — what the gradient path of this run is."""
    if self.world == 1:
      return "none"
    return self.symm.collective() if self.fused else "nccl"

  def _results(self, sp: Optional[StepPlan], with_errors: bool, reduced: bool = True) -> Dict[str, torch.Tensor]:
    s = (self.symm.scalars_sum if (self.fused and reduced) else self.slots).clone()
    r = OrderedDict([
        ("loss_G/loss", s[0]), ("loss_G/cycle", s[1]), ("loss_G/identity", s[2]),
        ("loss_G/total", s[0:3].sum()),
        ("loss_F/loss", s[3]), ("loss_F/cycle", s[4]), ("loss_F/identity", s[5]),
        ("loss_F/total", s[3:6].sum()),
        ("loss_X/loss", s[6]), ("loss_Y/loss", s[7])])
    if with_errors:
      r["error/MAE(X, F(G(X)))"] = s[8]
      r["error/MAE(Y, G(F(Y)))"] = s[9]
      r["error/MAE(X, F(X))"] = s[10]
      r["error/MAE(Y, G(Y))"] = s[11]
    return r

  def _load_inputs(self, sp: StepPlan, x: torch.Tensor, y: torch.Tensor):
    sp.x_in.copy_(x.to(torch.float32), non_blocking=True)
    sp.y_in.copy_(y.to(torch.float32), non_blocking=True)

  def _zero_results(self, with_errors: bool):
    z = torch.zeros((), device=self.device)
    keys = ["loss_G/loss", "loss_G/cycle", "loss_G/identity", "loss_G/total", "loss_F/loss",
            "loss_F/cycle", "loss_F/identity", "loss_F/total", "loss_X/loss", "loss_Y/loss"]
    if with_errors:
      keys += ["error/MAE(X, F(G(X)))", "error/MAE(Y, G(F(Y)))", "error/MAE(X, F(X))", "error/MAE(Y, G(Y))"]
    return OrderedDict((k, z.clone()) for k in keys)

  # ----------------------------------------------------------------- steps
  def train_step(self, x: torch.Tensor, y: torch.Tensor, debug: bool = False) -> Dict[str, torch.Tensor]:
    b = x.shape[0]
    if b == 0:
      return self._empty_train_step()
    sp = self.plan_for(b)
    self._load_inputs(sp, x, y)
    end = sp.marks["end"]
    if self.world > 1 and not self.fused:
      self._run_segmented(sp)
    elif self.use_graph and not debug:
      if sp.graph is None:
        # the first step runs eagerly (also sets lazy kernel attributes); the graph is captured
        # right after it — capture records the launches without executing them
        sp.plan.run(0, end, self._stream(), False)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
          sp.plan.run(0, end, self._stream(), False)
        sp.graph = g
      else:
        sp.graph.replay()
    else:
      sp.plan.run(0, end, self._stream(), debug)
    self._launches += sp.kernels_per_step
    return self._results(sp, with_errors=False)

  def _run_segmented(self, sp: StepPlan):
    """NCCL baseline path: plan segments with one packed all-reduce per network in between."""
    s = self._stream()
    pos = 0
    for n in NETS:
      sp.plan.run(pos, sp.marks[f"{n}_bwd_end"], s, False)
      self.strategy.all_reduce_sum(self.state[n].g)
      pos = sp.marks[f"{n}_bwd_end"]
    sp.plan.run(pos, sp.marks["end"], s, False)

  def _empty_train_step(self):
    """A replica with an empty share of a partial batch still joins the collectives."""
    s = self._stream()
    if self.fused:
      # replay exactly the cross-GPU launch sequence of the step plan (comm_order), on one stream:
      # local gradients and loss slots are zero, the results are the sums over the other replicas
      if self._opt_only_plan is None:
        plan = nv.load().Plan()
        plan.add_memset("zero.slots", self.slots.data_ptr(), NUM_SLOTS * 4)
        for n, k in self.comm_order():
          self.emit_optimizer(plan, n, bucket=k)
        self._opt_only_plan = plan
      self._opt_only_plan.run(0, -1, s, False)
      return self._results(None, with_errors=False)
    for n in NETS:
      if self.world > 1:
        self.strategy.all_reduce_sum(self.state[n].g)
    # run just the optimizer ops of any existing plan (gradients are zero locally)
    sp = self.plan_for(max(self._plans.keys(), default=1))
    for n in NETS:
      sp.plan.run(sp.marks[f"{n}_bwd_end"], sp.marks[f"{n}_opt_end"], s, False)
    return self._zero_results(False)

  def _reduce_slots(self):
    """Fused path, steps without an optimizer launch (test step): stand-alone one-shot SUM of the
    loss slots over ranks (same peer-memory kernel, scalar side channel only)."""
    if self._scalar_plan is None:
      plan = nv.load().Plan()
      self.symm.emit_scalar_reduce(plan)
      self._scalar_plan = plan
    self._scalar_plan.run(0, -1, self._stream(), False)

  @torch.no_grad()
  def test_step(self, x: torch.Tensor, y: torch.Tensor) -> Dict[str, torch.Tensor]:
    b = x.shape[0]
    if b == 0:
      if not self.fused:
        return self._zero_results(True)
      self.slots.zero_()
    else:
      sp = self.plan_for(b)
      self._load_inputs(sp, x, y)
      sp.plan.run(0, sp.marks["test_end"], self._stream(), False)
    if self.fused:
      self._reduce_slots()
    return self._results(None, with_errors=True)

  @torch.no_grad()
  def cycle_step(self, x: torch.Tensor, y: torch.Tensor):
    """(fake_x, fake_y, cycle_x, cycle_y) as fp32 NHWC3 tensors."""
    b = x.shape[0]
    sp = self.plan_for(b)
    self._load_inputs(sp, x.to(self.device), y.to(self.device))
    sp.plan.run(0, sp.marks["cycle_end"], self._stream(), False)
    G, F = sp.ex["G"], sp.ex["F"]
    img = lambda e, s0: e.inp[-1].t[s0:s0 + b, :, :, :3].float()
    return img(F, 0), img(G, 0), img(F, 2 * b), img(G, 2 * b)

  # ------------------------------------------------------------ checkpoint
  def _full(self, n: str, t: torch.Tensor) -> torch.Tensor:
    """fp32 master / moments are sharded across ranks on the fused path (ZeRO-1)."""
    if self.fused:
      return self.symm.gather_sharded(t, n)
    return t

  def state_dict(self) -> Dict[str, torch.Tensor]:
    sd: Dict[str, torch.Tensor] = OrderedDict()
    for n in NETS:
      st = self.state[n]
      w = st.layout.unpack(self._full(n, st.w))
      m = st.layout.unpack(self._full(n, st.m))
      v = st.layout.unpack(self._full(n, st.v))
      for name, t in w.items():
        sd[f"{n}/{name}/.ATTRIBUTES/VARIABLE_VALUE"] = t
      hyper = {"iter": float(st.step.item()), "learning_rate": ADAM["lr"], "beta_1": ADAM["b1"],
               "beta_2": ADAM["b2"], "decay": 0.0}
      for hk, hv in hyper.items():
        sd[f"{n}_optimizer/{hk}/.ATTRIBUTES/VARIABLE_VALUE"] = torch.tensor(hv, dtype=torch.float64)
      for name in w:
        sd[f"{n}/{name}/.OPTIMIZER_SLOT/{n}_optimizer/m/.ATTRIBUTES/VARIABLE_VALUE"] = m[name]
        sd[f"{n}/{name}/.OPTIMIZER_SLOT/{n}_optimizer/v/.ATTRIBUTES/VARIABLE_VALUE"] = v[name]
    return sd

  def load_state_dict(self, sd: Dict[str, torch.Tensor], expect_partial: bool = False) -> None:
    """Restore weights + optimizer state. Like the reference's ``checkpoint.read()`` (main.py:162-170)
    missing OPTIMIZER state is not an error: absent Adam slots / counters are (re)initialised to zero
    with a warning (e.g. a bundle written before the first step, or weights-only exports). Missing
    WEIGHTS raise unless ``expect_partial``."""
    missing_w: List[str] = []
    missing_opt: List[str] = []
    for n in NETS:
      st = self.state[n]
      cur_w = st.layout.unpack(self._full(n, st.w))
      cur_m = {k: torch.zeros_like(v) for k, v in cur_w.items()}
      cur_v = {k: torch.zeros_like(v) for k, v in cur_w.items()}
      for name in cur_w:
        for store, key, miss in (
            (cur_w, f"{n}/{name}/.ATTRIBUTES/VARIABLE_VALUE", missing_w),
            (cur_m, f"{n}/{name}/.OPTIMIZER_SLOT/{n}_optimizer/m/.ATTRIBUTES/VARIABLE_VALUE", missing_opt),
            (cur_v, f"{n}/{name}/.OPTIMIZER_SLOT/{n}_optimizer/v/.ATTRIBUTES/VARIABLE_VALUE", missing_opt)):
          if key in sd:
            store[name] = sd[key].float()
          else:
            miss.append(key)
      st.w.copy_(st.layout.pack(cur_w).to(self.device))
      st.m.copy_(st.layout.pack(cur_m).to(self.device))
      st.v.copy_(st.layout.pack(cur_v).to(self.device))
      st.wb.copy_(st.w.to(torch.bfloat16))
      st.refresh_derived_now()
      key = f"{n}_optimizer/iter/.ATTRIBUTES/VARIABLE_VALUE"
      if key in sd:
        st.step.fill_(int(sd[key]))
      else:
        st.step.zero_()
        missing_opt.append(key)
    torch.cuda.synchronize(self.device)
    self.strategy.barrier()
    if missing_opt and self.strategy.is_chief:
      print(f"[checkpoint] {len(missing_opt)} optimizer entries missing (e.g. {missing_opt[0]}): "
            "Adam state for them starts from zero")
    if missing_w and not expect_partial:
      raise KeyError(f"checkpoint is missing {len(missing_w)} weight tensors, e.g. {missing_w[:3]} "
                     f"(pass expect_partial=True to ignore)")

  # FP8 delayed-scaling state rides in the checkpoint's ``extra`` side-car, so the first step after a
  # resume quantises with the scales it would have used without the interruption
  def extra_state(self) -> Dict[str, object]:
    out: Dict[str, object] = {}
    if self.fp8 == "tensor":
      out["fp8"] = {n: {"act_amax": self.state[n].act_amax.tolist(), "act_scale": self.state[n].act_scale.tolist(),
                        "w_scale": self.state[n].w_scale.tolist(), "out_scale": self.state[n].out_scale.tolist()}
                    for n in ("G", "F") if self.state[n].w8 is not None}
    return out

  def load_extra_state(self, extra: Dict[str, object]) -> None:
    fp8 = (extra or {}).get("fp8")
    if not (self.fp8 == "tensor" and fp8):
      return
    for n, d in fp8.items():
      st = self.state[n]
      if st.w8 is None:
        continue
      for k, v in d.items():
        t = getattr(st, k)
        if len(v) == t.numel():
          t.copy_(torch.tensor(v, dtype=t.dtype))
