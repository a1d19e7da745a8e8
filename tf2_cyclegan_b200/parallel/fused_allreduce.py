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

import os
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from .strategy import Strategy

NUM_BLOCKS = int(os.environ.get("B2_ALLREDUCE_BLOCKS", "24"))


class SymmetricBuffers:

  def __init__(self, strategy: Strategy):
    import torch.distributed._symmetric_memory as symm_mem
    self.symm_mem = symm_mem
    self.strategy = strategy
    self.rank, self.world = strategy.rank, strategy.world_size
    if self.world > 8:
      raise RuntimeError("fused all-reduce supports up to 8 ranks (one NVSwitch domain)")
    need = 2 * NUM_BLOCKS * self.world * 4
    if symm_mem.get_signal_pad_size() < need:
      symm_mem.set_signal_pad_size(max(need, 16384))
    self.handles: Dict[str, Tuple[object, object]] = {}
    self.buckets: Dict[str, List[Tuple[int, int]]] = {}
    self.use_multicast = os.environ.get("B2_NVLS", "auto")

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

  def set_buckets(self, name: str, bounds):
    """Split network ``name``'s flat buffer into all-reduce buckets [lo, hi) (multiples of 8
    elements, ascending, covering [0, n)). Every bucket is reduced / updated by its own kernel
    launch as soon as its gradients are final, and is sharded over the ranks on its own — so the
    optimizer-state ownership map (ZeRO-1) is per bucket: see :meth:`owned_ranges`."""
    bounds = [(int(lo), int(hi)) for lo, hi in bounds]
    assert all(lo % 8 == 0 and hi % 8 == 0 and lo < hi for lo, hi in bounds), bounds
    assert bounds[0][0] == 0 and all(a[1] == b[0] for a, b in zip(bounds, bounds[1:])), bounds
    self.buckets[name] = bounds

  def bucket_list(self, name: str, n: int):
    return self.buckets.get(name) or [(0, n)]

  def emit_allreduce_adam(self, plan, name: str, st, lr: float, b1: float, b2: float, eps: float,
                          bucket: Optional[int] = None):
    """One fused reduce + Adam + broadcast launch over the whole buffer (``bucket`` None: all
    buckets, back to back) or over one bucket (pointers offset to the bucket start)."""
    hg, hw = self.handles[name]
    mg, mw = self.multicast_ptrs(name)
    todo = self.bucket_list(name, st.n)
    if bucket is not None:
      todo = [todo[bucket]]
    for lo, hi in todo:
      plan.add_allreduce_adam(f"opt.{name}.fused_allreduce_adam[{lo}:{hi}]",
                              [int(p) + 4 * lo for p in hg.buffer_ptrs],
                              [int(p) + 2 * lo for p in hw.buffer_ptrs], [int(p) for p in hg.signal_pad_ptrs],
                              st.w.data_ptr() + 4 * lo, st.m.data_ptr() + 4 * lo, st.v.data_ptr() + 4 * lo,
                              st.step.data_ptr(), lr, b1, b2, eps, hi - lo, self.rank, self.world,
                              NUM_BLOCKS, mg + 4 * lo if mg else 0, mw + 2 * lo if mw else 0)

  def shard_bounds(self, n: int, rank: Optional[int] = None) -> Tuple[int, int]:
    """This rank's shard of a buffer (or bucket) of n elements — mirrors the kernel."""
    rank = self.rank if rank is None else rank
    n8 = n // 8
    per = (n8 + self.world - 1) // self.world
    lo = min(n8, per * rank)
    return lo * 8, min(n8, lo + per) * 8

  def owned_ranges(self, name: str, n: int):
    """[lo, hi) ranges of network ``name``'s flat buffers whose fp32 master / moments live on
    this rank: its shard of every bucket."""
    out = []
    for lo, hi in self.bucket_list(name, n):
      a, b = self.shard_bounds(hi - lo)
      if b > a:
        out.append((lo + a, lo + b))
    return out

  def gather_sharded(self, t: torch.Tensor, name: Optional[str] = None) -> torch.Tensor:
    """Assemble a full tensor from the per-rank live shards (ZeRO-1 layout)."""
    full = torch.zeros_like(t)
    ranges = self.owned_ranges(name, t.numel()) if name is not None else [self.shard_bounds(t.numel())]
    for lo, hi in ranges:
      full[lo:hi] = t[lo:hi]
    self.strategy.all_reduce_sum(full)
    return full
