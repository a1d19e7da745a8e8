"""Checkpoint container: same file names, cadence and semantics as the
reference's ``tf.train.Checkpoint.write(prefix)`` (main.py:157-170), own format.

``<prefix>.index``                  JSON manifest (key → dtype/shape/offset) + extras
``<prefix>.data-00000-of-00001``    raw little-endian tensor bytes, 64-byte aligned

Keys are the Keras object-graph names
(``G/layer_with_weights-0/kernel/.ATTRIBUTES/VARIABLE_VALUE``, optimizer slots
``…/.OPTIMIZER_SLOT/G_optimizer/m/…``) and tensors are in TF layouts (HWIO
kernels), so converting to/from a TensorBundle is pure re-packaging.

``save(..., fmt="tensorbundle")`` writes the reference's own on-disk format instead (a TensorFlow
TensorBundle, see ``tensorbundle.py``), and ``load`` auto-detects which of the two containers
``<prefix>.index`` is — a checkpoint written by the reference can be dropped into
``<output_dir>/checkpoints/`` and resumed from.

Beyond the reference: writes are atomic (tmp + fsync + rename, with a nonce
tying index and data together — the reference can corrupt its only checkpoint
if it dies mid-write), and ``extra`` carries epoch / RNG state.
"""
from __future__ import annotations

import json
import os
import secrets
from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch

FORMAT = "b200-cyclegan-ckpt-v1"
DATA_SUFFIX = ".data-00000-of-00001"
_ALIGN = 64

_DTYPES = {
    torch.float32: "float32", torch.float64: "float64", torch.bfloat16: "bfloat16",
    torch.float16: "float16", torch.int64: "int64", torch.int32: "int32",
    torch.uint8: "uint8",
}
_FROM = {v: k for k, v in _DTYPES.items()}


def _to_bytes(t: torch.Tensor) -> bytes:
  t = t.detach().cpu().contiguous()
  if t.dtype == torch.bfloat16:
    return t.view(torch.int16).numpy().tobytes()
  return t.numpy().tobytes()


def exists(prefix: str) -> bool:
  """The reference resumes iff ``<prefix>.index`` exists (main.py:164)."""
  return os.path.exists(prefix + ".index")


def save(prefix: str, tensors: Dict[str, torch.Tensor],
         extra: Optional[Dict[str, Any]] = None, fmt: str = "native") -> None:
  if fmt == "tensorbundle":
    from . import tensorbundle as tb
    tb.write_bundle(prefix, tb.to_reference_dtypes(tensors), extra=extra)
    return
  if fmt != "native":
    raise ValueError(f"unknown checkpoint format {fmt!r}")
  os.makedirs(os.path.dirname(prefix) or ".", exist_ok=True)
  nonce = secrets.token_hex(8)
  manifest, offset = {}, 0
  data_tmp = prefix + DATA_SUFFIX + ".tmp"
  with open(data_tmp, "wb") as f:
    for key, t in tensors.items():
      raw = _to_bytes(t)
      pad = (-offset) % _ALIGN
      f.write(b"\0" * pad)
      offset += pad
      manifest[key] = {"dtype": _DTYPES[t.dtype], "shape": list(t.shape),
                       "offset": offset, "nbytes": len(raw)}
      f.write(raw)
      offset += len(raw)
    f.write(nonce.encode())
    f.flush()
    os.fsync(f.fileno())
  index_tmp = prefix + ".index.tmp"
  with open(index_tmp, "w") as f:
    json.dump({"format": FORMAT, "nonce": nonce, "data_bytes": offset,
               "extra": extra or {}, "tensors": manifest}, f)
    f.flush()
    os.fsync(f.fileno())
  os.replace(data_tmp, prefix + DATA_SUFFIX)
  os.replace(index_tmp, prefix + ".index")


def load(prefix: str) -> Tuple[Dict[str, torch.Tensor], Dict[str, Any]]:
  from . import tensorbundle as tb
  if tb.is_bundle(prefix):
    tensors, info = tb.read_bundle(prefix)
    return tb.from_reference_dtypes(tensors), info["extra"]
  with open(prefix + ".index") as f:
    index = json.load(f)
  if index.get("format") != FORMAT:
    raise ValueError(f"{prefix}.index: unknown checkpoint format {index.get('format')!r}")
  with open(prefix + DATA_SUFFIX, "rb") as f:
    blob = f.read()
  n = index["data_bytes"]
  if blob[n:n + 16].decode(errors="replace") != index["nonce"]:
    raise IOError(f"{prefix}: index and data files are from different writes "
                  "(interrupted save?)")
  out: Dict[str, torch.Tensor] = {}
  for key, meta in index["tensors"].items():
    raw = blob[meta["offset"]:meta["offset"] + meta["nbytes"]]
    dt = _FROM[meta["dtype"]]
    if dt == torch.bfloat16:
      arr = np.frombuffer(raw, dtype=np.int16).copy()
      t = torch.from_numpy(arr).view(torch.bfloat16)
    else:
      arr = np.frombuffer(raw, dtype=np.dtype(meta["dtype"])).copy()
      t = torch.from_numpy(arr)
    out[key] = t.reshape(meta["shape"])
  return out, index.get("extra", {})
