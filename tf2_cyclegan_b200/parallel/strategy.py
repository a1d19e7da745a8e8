"""Data-parallel strategy: the B200-native counterpart of
``tf.distribute.MirroredStrategy`` (main.py:370).

The reference drives all GPUs from one process; here it is one process per
GPU (``torchrun``-style: RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the
environment), which is what CUDA-graph capture and per-rank device timing
need. ``torch.distributed`` (NCCL on GPUs, gloo on CPU) provides bootstrap,
barriers and the *baseline* collectives; the product path for the gradient
all-reduce is the fused peer-memory kernel in parallel/fused_allreduce.py.

Semantics kept from the reference:
  * ``num_replicas_in_sync``; global batch = replicas × --batch_size (main.py:371-372)
  * each global batch is split across replicas as evenly as possible, including
    the partial last batch (``experimental_distribute_dataset``, main.py:80-81)
  * scalar results are SUM-reduced across replicas (main.py:264-267)
"""
from __future__ import annotations

import datetime
import os
from typing import Dict, Optional, Sequence

import torch
import torch.distributed as dist


class Strategy:

  def __init__(self, device: Optional[str] = None, backend: Optional[str] = None,
               init_process_group: bool = True):
    self.rank = int(os.environ.get("RANK", "0"))
    self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if device is None:
      device = "cuda" if torch.cuda.is_available() else "cpu"
    if device == "cuda":
      torch.cuda.set_device(self.local_rank)
      self.device = torch.device("cuda", self.local_rank)
    else:
      self.device = torch.device(device)
    self.group = None
    self._owns_group = False
    if self.world_size > 1 and init_process_group:
      if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if self.device.type == "cuda" else "gloo")
        kwargs = {}
        if backend == "nccl":
          kwargs["device_id"] = self.device
        dist.init_process_group(backend=backend, rank=self.rank,
                                world_size=self.world_size,
                                timeout=datetime.timedelta(minutes=10), **kwargs)
        self._owns_group = True
      self.group = dist.group.WORLD

  # -- MirroredStrategy vocabulary
  @property
  def num_replicas_in_sync(self) -> int:
    return self.world_size

  @property
  def is_chief(self) -> bool:
    return self.rank == 0

  # -- collectives (baseline path)
  def all_reduce_sum(self, t: torch.Tensor) -> torch.Tensor:
    if self.world_size > 1:
      dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
    return t

  def all_reduce_max(self, t: torch.Tensor) -> torch.Tensor:
    if self.world_size > 1:
      dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
    return t

  def broadcast(self, t: torch.Tensor, src: int = 0) -> torch.Tensor:
    if self.world_size > 1:
      dist.broadcast(t, src=src, group=self.group)
    return t

  def barrier(self) -> None:
    if self.world_size > 1:
      if self.device.type == "cuda":
        dist.barrier(group=self.group, device_ids=[self.local_rank])
      else:
        dist.barrier(group=self.group)

  def reduce_dict(self, d: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """SUM every scalar across replicas with ONE packed collective
    (the reference issues one ``strategy.reduce`` per key, main.py:264-267)."""
    if self.world_size == 1 or not d:
      return d
    keys = list(d.keys())
    packed = torch.stack([d[k].detach().reshape(()).float() for k in keys])
    self.all_reduce_sum(packed)
    return {k: packed[i] for i, k in enumerate(keys)}

  # -- dataset sharding
  def shard_bounds(self, n: int) -> Sequence[int]:
    """[start, end) of this replica's share of a global batch of n samples.

    Splits as evenly as possible with the remainder going to the lowest ranks,
    so a partial last batch gives uneven (possibly empty) shares."""
    base, rem = divmod(n, self.world_size)
    start = self.rank * base + min(self.rank, rem)
    return start, start + base + (1 if self.rank < rem else 0)

  def shard(self, t: torch.Tensor) -> torch.Tensor:
    s, e = self.shard_bounds(t.shape[0])
    return t[s:e]

  def shutdown(self) -> None:
    if self._owns_group and dist.is_initialized():
      dist.destroy_process_group()
      self._owns_group = False
