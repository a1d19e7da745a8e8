"""ctypes binding of ``csrc/_host.so`` (crc32c, fused image preprocessing, batch gather).

The library is built by ``csrc/build.py`` (g++, no CUDA). Every entry point has a pure
Python / PyTorch fallback with identical results, so CPU-only installs work without it;
``available()`` tells which one is in use.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "csrc")
_lib: Optional[ctypes.CDLL] = None
_tried = False


def _load() -> Optional[ctypes.CDLL]:
  global _lib, _tried
  if _tried:
    return _lib
  _tried = True
  path = os.path.join(_HERE, "_host.so")
  if not os.path.exists(path) and os.environ.get("B2_NO_HOST_BUILD", "0") != "1":
    try:
      import importlib.util
      spec = importlib.util.spec_from_file_location("_b2_csrc_build", os.path.join(_HERE, "build.py"))
      mod = importlib.util.module_from_spec(spec)
      spec.loader.exec_module(mod)
      mod.build_host()
    except Exception:
      return None
  try:
    lib = ctypes.CDLL(path)
  except OSError:
    return None
  lib.b2h_crc32c.restype = ctypes.c_uint32
  lib.b2h_crc32c.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32]
  lib.b2h_preprocess_batch.restype = None
  lib.b2h_preprocess_batch.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int]
  lib.b2h_gather_rows.restype = None
  lib.b2h_gather_rows.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                  ctypes.c_size_t, ctypes.c_int]
  _lib = lib
  return lib


def available() -> bool:
  return _load() is not None


def default_threads() -> int:
  return max(1, min(16, (os.cpu_count() or 2) // 2))


# ------------------------------------------------------------------ crc32c
_TABLE = None


def _crc32c_py(data: bytes, crc: int = 0) -> int:
  global _TABLE
  if _TABLE is None:
    tbl = []
    for i in range(256):
      c = i
      for _ in range(8):
        c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
      tbl.append(c)
    _TABLE = tbl
  crc ^= 0xFFFFFFFF
  tbl = _TABLE
  for b in data:
    crc = tbl[(crc ^ b) & 0xFF] ^ (crc >> 8)
  return crc ^ 0xFFFFFFFF


def crc32c(data, crc: int = 0) -> int:
  """CRC-32C of a bytes-like object or a C-contiguous numpy array (continuing from ``crc``)."""
  lib = _load()
  if isinstance(data, np.ndarray):
    arr = np.ascontiguousarray(data)
    if lib is None:
      return _crc32c_py(arr.tobytes(), crc)
    return int(lib.b2h_crc32c(arr.ctypes.data, arr.nbytes, crc))
  data = bytes(data)
  if lib is None:
    return _crc32c_py(data, crc)
  return int(lib.b2h_crc32c(data, len(data), crc))


def masked_crc32c(data) -> int:
  """TFRecord / TensorBundle masking: rotate right by 15 and add a constant."""
  crc = crc32c(data)
  return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def unmask_crc32c(masked: int) -> int:
  rot = (masked - 0xA282EAD8) & 0xFFFFFFFF
  return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ------------------------------------------------------- image preprocessing
def preprocess_batch(images: np.ndarray, params: np.ndarray, resize_hw, crop_hw,
                     threads: Optional[int] = None, out: Optional[np.ndarray] = None) -> Optional[np.ndarray]:
  """normalise(crop(resize(flip(img)))) for uint8 ``images`` [n, H, W, C]; ``params`` is int32
  [n, 3] = (flip, top, left). Returns float32 [n, crop_h, crop_w, C], or None without the
  library (callers then use the PyTorch path). ``out``: optional preallocated destination."""
  lib = _load()
  if lib is None:
    return None
  images = np.ascontiguousarray(images, dtype=np.uint8)
  params = np.ascontiguousarray(params, dtype=np.int32)
  n, H, W, C = images.shape
  shape = (n, int(crop_hw[0]), int(crop_hw[1]), C)
  if out is None:
    out = np.empty(shape, dtype=np.float32)
  assert out.shape == shape and out.dtype == np.float32 and out.flags["C_CONTIGUOUS"]
  lib.b2h_preprocess_batch(images.ctypes.data, n, H, W, C, out.ctypes.data, params.ctypes.data,
                           int(resize_hw[0]), int(resize_hw[1]), int(crop_hw[0]), int(crop_hw[1]),
                           threads or default_threads())
  return out


def gather_rows(dst, src, index, threads: Optional[int] = None) -> bool:
  """dst[i] = src[index[i]] for contiguous CPU torch tensors; False without the library."""
  lib = _load()
  if lib is None:
    return False
  n = int(index.numel())
  if n == 0:
    return True
  assert src.is_contiguous() and dst.is_contiguous() and index.dtype.is_signed
  row_bytes = src[0].numel() * src.element_size()
  lib.b2h_gather_rows(dst.data_ptr(), src.data_ptr(), index.data_ptr(), n, row_bytes,
                      threads or min(4, default_threads()))
  return True
