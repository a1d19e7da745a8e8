"""csrc/hostlib.cpp (native host helpers) vs their Python / PyTorch definitions."""
import numpy as np
import pytest
import torch

from tf2_cyclegan_b200.data import pipeline as P
from tf2_cyclegan_b200.ops import hostlib

needs_lib = pytest.mark.skipif(not hostlib.available(), reason="_host.so not built")


def test_crc32c_known_vectors():
  # RFC 3720 B.4 test vectors
  assert hostlib.crc32c(b"123456789") == 0xE3069283
  assert hostlib.crc32c(bytes(32)) == 0x8A9136AA
  assert hostlib.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
  assert hostlib._crc32c_py(b"123456789") == 0xE3069283
  m = hostlib.masked_crc32c(b"abc")
  assert hostlib.unmask_crc32c(m) == hostlib.crc32c(b"abc")


@needs_lib
def test_crc32c_native_matches_python_and_chains():
  rng = np.random.default_rng(0)
  for n in (0, 1, 7, 8, 9, 63, 1000, 4099):
    a = rng.integers(0, 256, n, dtype=np.uint8)
    assert hostlib.crc32c(a) == hostlib._crc32c_py(a.tobytes())
    assert hostlib.crc32c(a[1:]) == hostlib._crc32c_py(a[1:].tobytes())   # unaligned start
  a = rng.integers(0, 256, 5000, dtype=np.uint8)
  assert hostlib.crc32c(a[1234:], hostlib.crc32c(a[:1234])) == hostlib.crc32c(a)


class _OddSource:
  def get(self, split, i):
    return np.random.default_rng(i).integers(0, 256, (300 + 7 * i, 317 - 5 * i, 3), dtype=np.uint8)


@needs_lib
@pytest.mark.parametrize("source", ["synthetic", "odd"])
def test_native_preprocess_matches_torch_path(source):
  src = P.SyntheticSource(raw_size=128, sizes={"trainA": 6}) if source == "synthetic" else _OddSource()
  resize, crop = (143, 143), 128
  r1, r2 = np.random.default_rng(3), np.random.default_rng(3)
  a = P._materialise_native(src, "trainA", 6, r1, resize, crop, chunk=4)
  b = P._materialise(src, "trainA", 6, lambda im, i: P.preprocess_train(im, r2, resize, crop))
  assert a.shape == b.shape == (6, crop, crop, 3)
  assert (a - b).abs().max() < 1e-4
  assert r1.random() == r2.random()          # both paths consumed the same random draws
  a = P._materialise_native(src, "trainA", 6, None, resize, crop)
  b = P._materialise(src, "trainA", 6, lambda im, i: P.preprocess_test(im, crop))
  assert (a - b).abs().max() < 1e-4


@needs_lib
def test_gather_rows():
  src = torch.randn(50, 5, 7, 3)
  idx = torch.tensor([3, 49, 0, 3, 17], dtype=torch.long)
  dst = torch.empty(5, 5, 7, 3)
  assert hostlib.gather_rows(dst, src, idx)
  assert torch.equal(dst, src[idx])


def test_pipeline_without_native_library_matches(monkeypatch):
  """Installs without csrc/_host.so use the PyTorch / Python paths: same datasets, same batches."""
  import types
  from tf2_cyclegan_b200.parallel.strategy import Strategy

  def datasets():
    args = types.SimpleNamespace(image_size=32, synthetic_size=6, seed=7, global_batch_size=2, data_dir=None)
    tr, te, pl = P.get_datasets(args, Strategy(device="cpu"))
    return [torch.cat([x, y]) for x, y in tr], [torch.cat([x, y]) for x, y in te]

  with_lib = datasets()
  monkeypatch.setattr(hostlib, "_tried", True)
  monkeypatch.setattr(hostlib, "_lib", None)
  assert not hostlib.available()
  without = datasets()
  for a, b in zip(with_lib[0] + with_lib[1], without[0] + without[1]):
    assert a.shape == b.shape and (a - b).abs().max() < 1e-4
  assert hostlib.crc32c(b"123456789") == 0xE3069283
