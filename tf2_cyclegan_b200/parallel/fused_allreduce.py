"""Symmetric (peer-mapped) buffers for the fused gradient reduce + Adam kernel.

Each network's fp32 gradient buffer and bf16 weight copy are allocated with
``torch.distributed._symmetric_memory`` (CUDA VMM + cross-process handle
exchange), which gives every rank the peers' device pointers, a zero-initialised
signal pad and — when the fabric supports NVLS — a multicast alias of the same
allocation. Those raw pointers are baked into ``csrc/allreduce_adam.cu``'s
parameter block; NCCL is used only for bootstrap and for checkpoint gathers.

Replaces: the four ``NcclAllReduce(num_packs=1)`` + 168 ``ResourceApplyAdam``
launches issued per step by the reference's ``optimizer.minimize`` calls
(main.py:249-260).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from .strategy import Strategy

MAX_BLOCKS = 148        # signal-pad sizing: one flag set per block and phase
NUM_SCALARS = 16        # loss / metric slots reduced through the side channel
from ..config import ENGINE

# Form of the two-shot part: "push" (posted peer stores into the owner's staging slots, then local
# reduction: bandwidth bound with a few CTAs) or "pull" (owner loads the peers' shards / NVLS
# multimem.ld_reduce: needs megabytes of loads in flight). See csrc/allreduce_adam.cuh; chosen by
# config.ENGINE.ar_mode.
DEFAULT_MODE = ENGINE.ar_mode


def default_blocks(nbytes: int, exposed: bool, world: int = 2, mode: Optional[str] = None) -> int:
  """CTAs of one fused reduce + Adam launch (always an even number: they are launched as cluster pairs,
  one TPC per pair). Launches that run under the backward pass use few — they displace GEMM CTA pairs one
  for one; the last, exposed launch of a network has the machine to itself.
  Pull form (512 threads, 64 KB of peer loads in flight per CTA): one CTA per 128 KB of the rank's shard,
  2..12, exposed: per 32 KB up to 148 — an in-switch reduction over 8 sources has a ~20 us round trip.
  Push form (1024 threads, 128 KB of local loads in flight): one CTA per 2 MB, 2..8, exposed per 512 KB
  up to 32."""
  mode = mode or DEFAULT_MODE
  if mode == "push":
    want = max(2, min(8, -(-nbytes // (2 << 20))))
    if exposed:
      want = max(want, min(32, -(-nbytes // (512 << 10))))
  else:
    shard = max(1, nbytes // max(1, world))
    want = max(2, min(12, -(-shard // (128 * 1024))))
    if exposed:
      want = min(MAX_BLOCKS, max(want, -(-shard // (32 * 1024))))
  return want + (want & 1)


class SymmetricBuffers:

  def __init__(self, strategy: Strategy):
    import torch.distributed._symmetric_memory as symm_mem
    self.symm_mem = symm_mem
    self.strategy = strategy
    self.rank, self.world = strategy.rank, strategy.world_size
    if self.world > 8:
      raise RuntimeError("fused all-reduce supports up to 8 ranks (one NVSwitch domain)")
    need = 2 * MAX_BLOCKS * self.world * 4
    if symm_mem.get_signal_pad_size() < need:
      symm_mem.set_signal_pad_size(max(need, 16384))
    self.handles: Dict[str, Tuple[object, object]] = {}
    self.buckets: Dict[str, List[Tuple[int, int]]] = {}
    self.vec_lo: Dict[str, int] = {}
    self.use_multicast = ENGINE.nvls
    # loss-scalar side channel: every rank's slots are peer-readable; the sums land in `scalars_sum`
    self.scalars = symm_mem.empty(NUM_SCALARS, dtype=torch.float32, device=strategy.device)
    self.scalars.zero_()
    self.h_scalars = symm_mem.rendezvous(self.scalars, dist.group.WORLD)
    self.scalars_sum = torch.zeros(NUM_SCALARS, device=strategy.device)
    self._dummy: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
    self._stages: List[Tuple[torch.Tensor, object]] = []     # push-mode staging buffers (grow-only)

  def stage_ptrs(self, n: int) -> List[int]:
    """Peer pointers of a staging buffer with room for a two-shot region of n elements:
    [world][ceil(n / 8 / world) * 8] floats. Collective (symmetric allocation) when it has to grow;
    launches that were planned against an older, smaller buffer keep it alive."""
    need = self.world * (-(-(n // 8) // self.world)) * 8
    if not self._stages or self._stages[-1][0].numel() < need:
      t = self.symm_mem.empty(max(need, 1 << 20), dtype=torch.float32, device=self.strategy.device)
      t.zero_()
      self._stages.append((t, self.symm_mem.rendezvous(t, dist.group.WORLD)))
    return [int(p) for p in self._stages[-1][1].buffer_ptrs]

  def alloc(self, name: str, numel: int):
    sm = self.symm_mem
    dev = self.strategy.device
    g = sm.empty(numel, dtype=torch.float32, device=dev)
    wb = sm.empty(numel, dtype=torch.bfloat16, device=dev)
    g.zero_()
    wb.zero_()
    hg = sm.rendezvous(g, dist.group.WORLD)
    hw = sm.rendezvous(wb, dist.group.WORLD)
    self.handles[name] = (hg, hw)
    return g, wb

  def collective(self) -> str:
    """What the gradient path of this run is: 'fused-<push|pull>-<nvls|p2p>' (nvls: the weight
    broadcast — and, in pull form, the reduction — goes through the NVSwitch multicast object)."""
    for name in self.handles:
      return f"fused-{DEFAULT_MODE}-" + ("nvls" if self.multicast_ptrs(name)[0] else "p2p")
    return f"fused-{DEFAULT_MODE}-p2p"

  def multicast_ptrs(self, name: str) -> Tuple[int, int]:
    hg, hw = self.handles[name]
    if self.use_multicast in ("0", "off"):
      return 0, 0
    try:
      mg, mw = int(hg.multicast_ptr), int(hw.multicast_ptr)
    except Exception:
      return 0, 0
    if mg == 0 or mw == 0:
      return 0, 0
    return mg, mw

  def set_buckets(self, name: str, bounds, vec_lo: Optional[int] = None):
    """Split the two-shot part [0, vec_lo) of network ``name``'s flat buffer into all-reduce buckets
    [lo, hi) (multiples of 8 elements, ascending, covering it). Every bucket is reduced / updated
    by its own kernel launch as soon as its gradients are final, and is sharded over the ranks on
    its own — so the optimizer-state ownership map (ZeRO-1) is per bucket: see :meth:`owned_ranges`.
    [vec_lo, n) is the one-shot (replicated) tail."""
    bounds = [(int(lo), int(hi)) for lo, hi in bounds]
    assert all(lo % 8 == 0 and hi % 8 == 0 and lo < hi for lo, hi in bounds), bounds
    assert bounds[0][0] == 0 and all(a[1] == b[0] for a, b in zip(bounds, bounds[1:])), bounds
    if vec_lo is not None:
      assert bounds[-1][1] == vec_lo and vec_lo % 8 == 0, (bounds, vec_lo)
      self.vec_lo[name] = int(vec_lo)
    self.buckets[name] = bounds

  def bucket_list(self, name: str, n: int):
    return self.buckets.get(name) or [(0, self.vec_lo.get(name, n))]

  def emit_allreduce_adam(self, plan, name: str, st, lr: float, b1: float, b2: float, eps: float,
                          bucket: Optional[int] = None, vectors: Optional[bool] = None,
                          scalars: bool = False, exposed: bool = False, mode: Optional[str] = None,
                          blocks: Optional[int] = None):
    """Fused reduce + Adam + broadcast launches of network ``name``: every bucket back to back
    (``bucket`` None) or ONE bucket. ``vectors``: also update the one-shot tail [vec_lo, n) in the
    (last) launch (default: when all buckets are emitted); ``scalars``: carry the loss-scalar sums."""
    hg, hw = self.handles[name]
    mg, mw = self.multicast_ptrs(name)
    todo = self.bucket_list(name, st.n)
    if vectors is None:
      vectors = bucket is None
    if bucket is not None:
      todo = [todo[bucket]]
    v_lo = self.vec_lo.get(name, st.n)
    mode = mode or DEFAULT_MODE
    assert mode in ("push", "pull"), mode
    for k, (lo, hi) in enumerate(todo):
      last = k == len(todo) - 1
      vec = (v_lo, st.n - v_lo) if (vectors and last) else (0, 0)
      sc = scalars and last
      plan.add_allreduce_adam(f"opt.{name}.fused_allreduce_adam[{lo}:{hi}]" + ("+vec" if vec[1] else "") +
                              ("+scalars" if sc else ""),
                              [int(p) for p in hg.buffer_ptrs], [int(p) for p in hw.buffer_ptrs],
                              [int(p) for p in hg.signal_pad_ptrs],
                              st.w.data_ptr(), st.m.data_ptr(), st.v.data_ptr(), st.step.data_ptr(),
                              lr, b1, b2, eps, lo, hi - lo, vec[0], vec[1],
                              [int(p) for p in self.h_scalars.buffer_ptrs] if sc else [],
                              self.scalars_sum.data_ptr() if sc else 0, NUM_SCALARS if sc else 0,
                              self.rank, self.world, blocks or default_blocks(4 * (hi - lo), exposed, self.world, mode),
                              mg, mw, self.stage_ptrs(hi - lo) if mode == "push" else [])

  def emit_scalar_reduce(self, plan, name: str = "scalars.allreduce"):
    """Stand-alone one-shot SUM of the loss slots (test step: no optimizer launch to ride on)."""
    if self._dummy is None:
      self._dummy = self.alloc("__scalars__", 128)
    hg, hw = self.handles["__scalars__"]
    # no parameter region at all (n = v_n = 0): master / moment pointers are never dereferenced
    plan.add_allreduce_adam(name, [int(p) for p in hg.buffer_ptrs], [int(p) for p in hw.buffer_ptrs],
                            [int(p) for p in hg.signal_pad_ptrs], 0, 0, 0,
                            self._zero_step().data_ptr(), 0.0, 0.5, 0.9, 1e-7, 0, 0, 0, 0,
                            [int(p) for p in self.h_scalars.buffer_ptrs], self.scalars_sum.data_ptr(),
                            NUM_SCALARS, self.rank, self.world, 1, 0, 0)

  def _zero_step(self) -> torch.Tensor:
    if not hasattr(self, "_zstep"):
      self._zstep = torch.zeros(1, dtype=torch.int32, device=self.strategy.device)
    return self._zstep

  def shard_bounds(self, n: int, rank: Optional[int] = None) -> Tuple[int, int]:
    """This rank's shard of a bucket of n elements — mirrors the kernel."""
    rank = self.rank if rank is None else rank
    n8 = n // 8
    per = (n8 + self.world - 1) // self.world
    lo = min(n8, per * rank)
    return lo * 8, min(n8, lo + per) * 8

  def owned_ranges(self, name: str, n: int):
    """[lo, hi) ranges of network ``name``'s flat buffers whose fp32 master / moments this rank
    contributes to a gather: its shard of every bucket; the replicated tail is taken from rank 0."""
    out = []
    for lo, hi in self.bucket_list(name, n):
      a, b = self.shard_bounds(hi - lo)
      if b > a:
        out.append((lo + a, lo + b))
    v_lo = self.vec_lo.get(name, n) if name is not None else n
    if v_lo < n and self.rank == 0:
      out.append((v_lo, n))
    return out

  def gather_sharded(self, t: torch.Tensor, name: Optional[str] = None) -> torch.Tensor:
    """Assemble a full tensor from the per-rank live shards (ZeRO-1 layout)."""
    full = torch.zeros_like(t)
    ranges = self.owned_ranges(name, t.numel()) if name is not None else [self.shard_bounds(t.numel())]
    for lo, hi in ranges:
      full[lo:hi] = t[lo:hi]
    self.strategy.all_reduce_sum(full)
    return full
