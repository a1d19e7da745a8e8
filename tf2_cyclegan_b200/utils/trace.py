"""NVTX ranges for timeline tools (the reference has no tracing at all).

``nvtx_range("train_step")`` is a no-op unless CUDA is available. The native plan adds one range
per op (label = the op name, e.g. ``G.res3a.fwd``) when ``config.ENGINE.nvtx`` (``B2_CONFIG=nvtx=1``) or
``Plan.set_nvtx(True)``; run with ``B2_CONFIG=use_graph=0`` so that the per-op ranges are visible (a
replayed CUDA graph is a single launch).
"""
from __future__ import annotations

import contextlib

import torch


@contextlib.contextmanager
def nvtx_range(name: str):
  on = torch.cuda.is_available()
  if on:
    torch.cuda.nvtx.range_push(name)
  try:
    yield
  finally:
    if on:
      torch.cuda.nvtx.range_pop()
