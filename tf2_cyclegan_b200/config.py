"""The ONE place where engine behaviour other than the CLI flags is configured.

``ENGINE`` is a plain dataclass instance read by the native backend, the planner and (through
``_C.set_kernel_options``) the C++ launchers. It can be changed

* programmatically:  ``from tf2_cyclegan_b200.config import ENGINE; ENGINE.use_graph = False``
  (before the backend / plans are built), or
* for one process from the shell through a single variable:
  ``B2_CONFIG="use_graph=0,lanes=2,ar_mode=pull" python bench.py ...``

Nothing else in the package reads the environment (``RANK`` / ``WORLD_SIZE`` / ``MASTER_*`` of
torchrun and the compilers' ``NVCC`` / ``CXX`` aside). The debugging fields are for bisecting
performance and make results invalid where noted.
"""
from __future__ import annotations

import dataclasses
import os
from dataclasses import dataclass


@dataclass
class EngineConfig:
  # ---- execution
  use_graph: bool = True       # replay the step plan as one CUDA graph (False: eager launches, e.g. under ncu)
  lanes: int = 4               # 1: one stream | 2: one lane per network pair | 4: + weight-gradient lanes
  nvtx: bool = False           # NVTX range per plan op (eager plans only)
  # ---- kernel selection
  tap_share: bool = True       # stride-1 N=256 convs: one halo tile per channel block (conv_gemm_ts.cu)
  two_cta: bool = True         # cta_group::2 kernels for wide layers
  wgrad_multi: bool = True     # narrow layers: all taps of a wgrad in one CTA (wgrad_multi.cu)
  wgrad_ctas: int = 0          # split-K target of the wgrad launches (0: whole-wave heuristic)
  # ---- data parallelism
  nvls: str = "auto"           # "auto": use the NVSwitch multicast alias when the fabric offers it | "off"
  ar_mode: str = "pull"        # two-shot part of the fused gradient reduce: "pull" | "push" (csrc/allreduce_adam.cuh);
                               # pull measured faster at 4 and 8 GPUs, in the step and stand-alone (profiles/README.md)
  # ---- debugging (timing experiments; results are INVALID for kernel_debug != 0)
  kernel_debug: int = 0        # conv kernels: 1 skip epilogue, 2 skip MMAs, 3 skip TMA loads, 4 skip global stores,
                               # 5 (MXFP8 kernels) skip the scale-factor stores to TMEM
  stages: int = 0              # 1-CTA conv kernel: cap on the smem ring depth (0: default)
  grid: int = 0                # 1-CTA conv kernel: cap on the persistent grid (0: all SMs)

  def update_from_string(self, spec: str) -> "EngineConfig":
    """``key=value[,key=value...]``; booleans accept 0/1/true/false."""
    fields = {f.name: f.type for f in dataclasses.fields(self)}
    for item in filter(None, (s.strip() for s in spec.split(","))):
      if "=" not in item:
        raise ValueError(f"B2_CONFIG: expected key=value, got {item!r}")
      k, v = (t.strip() for t in item.split("=", 1))
      if k not in fields:
        raise ValueError(f"B2_CONFIG: unknown key {k!r} (known: {', '.join(fields)})")
      cur = getattr(self, k)
      if isinstance(cur, bool):
        if v.lower() not in ("0", "1", "true", "false"):
          raise ValueError(f"B2_CONFIG: {k} is a boolean, got {v!r}")
        setattr(self, k, v.lower() in ("1", "true"))
      elif isinstance(cur, int):
        setattr(self, k, int(v))
      else:
        setattr(self, k, v)
    self.validate()
    return self

  def validate(self) -> None:
    if self.lanes not in (1, 2, 4):
      raise ValueError("config: lanes must be 1, 2 or 4")
    if self.ar_mode not in ("push", "pull"):
      raise ValueError("config: ar_mode must be 'push' or 'pull'")
    if self.nvls not in ("auto", "off"):
      raise ValueError("config: nvls must be 'auto' or 'off'")
    if not 0 <= self.kernel_debug <= 5:
      raise ValueError("config: kernel_debug must be 0..5")

  def apply_to_extension(self, mod) -> None:
    """Push the kernel options into the loaded C++ extension (no getenv on that side)."""
    mod.set_kernel_options(int(self.kernel_debug), bool(self.two_cta), int(self.stages), int(self.grid))


ENGINE = EngineConfig().update_from_string(os.environ.get("B2_CONFIG", ""))
