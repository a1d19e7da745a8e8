"""Input pipeline with the reference's tf.data semantics (main.py:18-83).

    take(n) → map(flip → bilinear resize 286² → random crop 256² → /127.5 − 1)
            → cache() → shuffle(256)            per train domain
    take(n) → map(resize 256² → normalise) → cache()        per test domain
    zip(A.batch(gbs), B.batch(gbs)) [.prefetch]             no drop_remainder
    distribute: every global batch is split across the replicas

Kept quirks: ``cache()`` sits *after* the random map, so flip/crop are drawn
once and frozen for the run; each domain is shuffled independently with a
256-element buffer (unpaired); the last batch is partial and the replicas get
uneven shares of it.

Sources: a TFDS-prepared ``cycle_gan/horse2zebra`` directory (the reference's
source; read without TensorFlow by data/tfds.py), a CycleGAN style folder
(``trainA/ trainB/ testA/ testB/``), or a deterministic synthetic horse2zebra
stand-in with the same split sizes (1067/1334/120/140).
The prefetcher stages batches in pinned host memory and copies them to the
device on a side stream (the B200 counterpart of MultiDeviceIterator).
"""
from __future__ import annotations

import atexit
import os
import threading
import queue
from math import ceil
from typing import Callable, Iterator, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from ..ops import hostlib
from ..parallel.strategy import Strategy

H2Z_SPLITS = {"trainA": 1067, "trainB": 1334, "testA": 120, "testB": 140}


# ------------------------------------------------------------------ sources
class SyntheticSource:
  """Procedural stand-in for cycle_gan/horse2zebra: domain A = smooth blobs,
  domain B = blobs modulated by stripes. Deterministic per (split, index)."""

  def __init__(self, raw_size: int = 256, sizes: Optional[dict] = None, seed: int = 1234):
    self.raw_size = raw_size
    self.sizes = dict(H2Z_SPLITS if sizes is None else sizes)
    self.seed = seed

  def num_examples(self, split: str) -> int:
    return self.sizes[split]

  def get(self, split: str, index: int) -> np.ndarray:
    s = self.raw_size
    rng = np.random.default_rng([self.seed, sum(map(ord, split)), index])
    low = rng.random((4, 4, 3), dtype=np.float32)
    img = F.interpolate(torch.from_numpy(low).permute(2, 0, 1)[None], size=(s, s),
                        mode="bicubic", align_corners=False)[0].permute(1, 2, 0).numpy()
    if split.endswith("B"):
      yy, xx = np.mgrid[0:s, 0:s].astype(np.float32) / s
      ang, freq = rng.random() * np.pi, 6 + 10 * rng.random()
      stripes = 0.5 + 0.5 * np.sign(np.sin(2 * np.pi * freq * (np.cos(ang) * xx + np.sin(ang) * yy)))
      img = img * (0.35 + 0.65 * stripes[..., None])
    return (np.clip(img, 0, 1) * 255).astype(np.uint8)


class FolderSource:
  """``<data_dir>/{trainA,trainB,testA,testB}/*.{jpg,png}`` (the layout of the
  original CycleGAN dataset archives)."""
  EXT = (".jpg", ".jpeg", ".png", ".bmp")

  def __init__(self, data_dir: str):
    self.files = {}
    for split in H2Z_SPLITS:
      d = os.path.join(data_dir, split)
      if not os.path.isdir(d):
        raise FileNotFoundError(f"missing split directory {d}")
      self.files[split] = sorted(os.path.join(d, f) for f in os.listdir(d)
                                 if f.lower().endswith(self.EXT))

  def num_examples(self, split: str) -> int:
    return len(self.files[split])

  def get(self, split: str, index: int) -> np.ndarray:
    from PIL import Image
    return np.asarray(Image.open(self.files[split][index]).convert("RGB"))


def open_source(data_dir: str):
  """``--data_dir`` auto-detection: a CycleGAN image folder (trainA/ ...) or a TFDS-prepared
  ``cycle_gan/horse2zebra`` directory (the reference's source, main.py:22-26)."""
  if all(os.path.isdir(os.path.join(data_dir, s)) for s in H2Z_SPLITS):
    return FolderSource(data_dir)
  from .tfds import TFDSSource, find_tfds_dir
  if find_tfds_dir(data_dir) is not None:
    return TFDSSource(data_dir)
  raise FileNotFoundError(f"{data_dir}: neither trainA/ trainB/ testA/ testB/ folders nor a TFDS-prepared "
                          f"cycle_gan/horse2zebra directory")


# ------------------------------------------------------------ preprocessing
def normalize_image(image: torch.Tensor) -> torch.Tensor:
  """normalize image to [-1, 1] (main.py:35-38)."""
  return image.to(torch.float32) / 127.5 - 1.0


def resize_bilinear(image: torch.Tensor, size: Tuple[int, int]) -> torch.Tensor:
  """tf.image.resize(method=bilinear): half-pixel centres, no antialias. HWC float."""
  x = image.to(torch.float32).permute(2, 0, 1)[None]
  y = F.interpolate(x, size=size, mode="bilinear", align_corners=False, antialias=False)
  return y[0].permute(1, 2, 0)


def preprocess_train(image: np.ndarray, rng: np.random.Generator,
                     resize_shape: Tuple[int, int], crop: int) -> torch.Tensor:
  """main.py:40-45: random flip → resize → random crop → normalise."""
  img = torch.from_numpy(np.ascontiguousarray(image))
  if rng.random() < 0.5:
    img = torch.flip(img, dims=[1])
  img = resize_bilinear(img, resize_shape)
  top = int(rng.integers(0, resize_shape[0] - crop + 1))
  left = int(rng.integers(0, resize_shape[1] - crop + 1))
  img = img[top:top + crop, left:left + crop, :]
  return normalize_image(img)


def preprocess_test(image: np.ndarray, size: int) -> torch.Tensor:
  """main.py:47-50: resize → normalise."""
  return normalize_image(resize_bilinear(torch.from_numpy(np.ascontiguousarray(image)), (size, size)))


def tf_shuffle_order(n: int, buffer_size: int, rng: np.random.Generator) -> List[int]:
  """Element order produced by ``Dataset.shuffle(buffer_size)`` over range(n)."""
  buf = list(range(min(buffer_size, n)))
  nxt = len(buf)
  out = []
  while buf:
    j = int(rng.integers(0, len(buf)))
    out.append(buf[j])
    if nxt < n:
      buf[j] = nxt
      nxt += 1
    else:
      buf[j] = buf[-1]
      buf.pop()
  return out


# ----------------------------------------------------------------- datasets
_LIVE_WORKERS: set = set()   # (stop event, thread) of prefetch workers whose consumer is still open


@atexit.register
def _join_prefetch_workers():
  """Iterators that are still open when the interpreter exits (a consumer that stopped early)
  must not leave their worker thread un-joined."""
  for stop, t in list(_LIVE_WORKERS):
    stop.set()
  for stop, t in list(_LIVE_WORKERS):
    t.join(timeout=10)
  _LIVE_WORKERS.clear()


class PairedBatches:
  """zip(A.batch(gbs), B.batch(gbs)) restricted to this replica's share."""

  def __init__(self, a: torch.Tensor, b: torch.Tensor, global_batch_size: int,
               strategy: Strategy, shuffle_buffer: int = 0, seed: int = 0,
               distribute: bool = True, prefetch: int = 0, to_device: bool = True):
    assert a.shape[0] == b.shape[0]
    self.a, self.b = a, b
    self.gbs = global_batch_size
    self.strategy = strategy
    self.shuffle_buffer = shuffle_buffer
    self.rng = np.random.default_rng(seed)
    self.distribute = distribute
    self.prefetch = prefetch
    self.to_device = to_device
    self.steps = ceil(a.shape[0] / global_batch_size)

  def __len__(self) -> int:
    return self.steps

  def _orders(self):
    n = self.a.shape[0]
    if self.shuffle_buffer:
      return (tf_shuffle_order(n, self.shuffle_buffer, self.rng),
              tf_shuffle_order(n, self.shuffle_buffer, self.rng))
    return list(range(n)), list(range(n))

  def _staging(self, k: int, n: int, like: torch.Tensor) -> torch.Tensor:
    """Rotating pinned host buffers (allocated once): the gather target of a batch and the
    source of its async H2D copy. `prefetch + 2` slots so a slot is never rewritten in flight."""
    dev = self.strategy.device
    if not self.to_device or dev.type != "cuda":
      return torch.empty((n,) + tuple(like.shape[1:]), dtype=like.dtype)
    if not hasattr(self, "_pinned"):
      self._pinned = {}
    key = (k, tuple(like.shape[1:]))
    if key not in self._pinned:
      cap = max(1, -(-self.gbs // self.strategy.world_size) if self.distribute else self.gbs)
      self._pinned[key] = torch.empty((cap,) + tuple(like.shape[1:]), dtype=like.dtype).pin_memory()
    return self._pinned[key][:n]

  def _host_batches(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
    oa, ob = self._orders()
    n = self.a.shape[0]
    slots = max(2, self.prefetch + 2)
    for step, s in enumerate(range(0, n, self.gbs)):
      ia, ib = oa[s:s + self.gbs], ob[s:s + self.gbs]
      if self.distribute:
        lo, hi = self.strategy.shard_bounds(len(ia))
        ia, ib = ia[lo:hi], ib[lo:hi]
      idx_a = torch.as_tensor(ia, dtype=torch.long)
      idx_b = torch.as_tensor(ib, dtype=torch.long)
      k = step % slots
      ev = getattr(self, "_slot_events", {}).get(k)
      if ev is not None:
        ev.synchronize()   # the async H2D copy that last read this pinned slot has finished
      self._slot = k
      xa = self._staging(2 * k, len(ia), self.a)
      xb = self._staging(2 * k + 1, len(ib), self.b)
      if len(ia):   # batch(): native multi-threaded row gather into the (pinned) staging slot
        if not (self.a.is_contiguous() and xa.is_contiguous()
                and hostlib.gather_rows(xa, self.a, idx_a)):
          torch.index_select(self.a, 0, idx_a, out=xa)
        if not (self.b.is_contiguous() and xb.is_contiguous()
                and hostlib.gather_rows(xb, self.b, idx_b)):
          torch.index_select(self.b, 0, idx_b, out=xb)
      yield xa, xb

  def _upload(self, x: torch.Tensor, y: torch.Tensor):
    dev = self.strategy.device
    if not self.to_device or dev.type != "cuda":
      return x, y
    return x.to(dev, non_blocking=True), y.to(dev, non_blocking=True)

  def __iter__(self):
    if self.prefetch <= 0:
      for x, y in self._host_batches():
        xd, yd = self._upload(x, y)
        if xd is not x:   # async copy from a pinned slot: finish it before the slot can be rewritten
          torch.cuda.current_stream(self.strategy.device).synchronize()
        yield xd, yd
      return
    q: "queue.Queue" = queue.Queue(maxsize=self.prefetch)
    dev = self.strategy.device
    stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" and self.to_device else None
    stop = threading.Event()

    def put(item) -> bool:
      while not stop.is_set():
        try:
          q.put(item, timeout=0.1)
          return True
        except queue.Full:
          continue
      return False

    def worker():
      try:
        for x, y in self._host_batches():
          if stream is not None:
            with torch.cuda.stream(stream):
              xd, yd = self._upload(x, y)
              ev = torch.cuda.Event()
              ev.record(stream)
            if not hasattr(self, "_slot_events"):
              self._slot_events = {}
            self._slot_events[self._slot] = ev
            ok = put((xd, yd, ev))
          else:
            ok = put((x, y, None))
          if not ok:
            return
      except BaseException as e:  # surface worker errors in the consumer
        put(e)
      put(None)

    t = threading.Thread(target=worker, daemon=True)
    t.start()
    _LIVE_WORKERS.add((stop, t))
    try:
      while True:
        item = q.get()
        if item is None:
          break
        if isinstance(item, BaseException):
          raise item
        x, y, ev = item
        if ev is not None:
          torch.cuda.current_stream(dev).wait_event(ev)
          x.record_stream(torch.cuda.current_stream(dev))
          y.record_stream(torch.cuda.current_stream(dev))
        yield x, y
    finally:
      # also runs when the consumer stops early (generator closed / garbage collected): the
      # prefetch thread must be joined, an un-joined thread that used the intra-op pool aborts
      # the interpreter at exit
      stop.set()
      t.join(timeout=10)
      _LIVE_WORKERS.discard((stop, t))


def _materialise_native(source, split: str, n: int, rng: Optional[np.random.Generator],
                        resize: Tuple[int, int], crop: int, chunk: int = 64) -> Optional[torch.Tensor]:
  """The per-image map of a split through ``csrc/hostlib.cpp`` (fused flip / bilinear resize /
  crop / normalise on a thread pool). ``rng`` = None: test map (resize straight to ``crop``).
  Random draws are consumed in the same order as :func:`preprocess_train` (flip, top, left per
  image), so both paths pick identical augmentations. Returns None without the library."""
  if not hostlib.available() or n == 0:
    return None
  out = None
  for s0 in range(0, n, chunk):
    imgs = [np.ascontiguousarray(source.get(split, i)) for i in range(s0, min(n, s0 + chunk))]
    params = np.zeros((len(imgs), 3), dtype=np.int32)
    if rng is not None:
      for j in range(len(imgs)):
        params[j, 0] = 1 if rng.random() < 0.5 else 0
        params[j, 1] = int(rng.integers(0, resize[0] - crop + 1))
        params[j, 2] = int(rng.integers(0, resize[1] - crop + 1))
    rs = resize if rng is not None else (crop, crop)
    if out is None:
      out = torch.empty((n, crop, crop, imgs[0].shape[2]), dtype=torch.float32)
    dst = out.numpy()[s0:s0 + len(imgs)]   # written in place by the library
    if all(im.shape == imgs[0].shape for im in imgs):
      hostlib.preprocess_batch(np.stack(imgs), params, rs, (crop, crop), out=dst)
    else:   # mixed raw sizes: one call per image
      for j, im in enumerate(imgs):
        hostlib.preprocess_batch(im[None], params[j:j + 1], rs, (crop, crop), out=dst[j:j + 1])
  return out


def _materialise(source, split: str, n: int, fn: Callable[[np.ndarray, int], torch.Tensor]) -> torch.Tensor:
  first = fn(source.get(split, 0), 0)
  out = torch.empty((n,) + tuple(first.shape), dtype=torch.float32)
  out[0] = first
  for i in range(1, n):
    out[i] = fn(source.get(split, i), i)
  return out


def get_datasets(args, strategy: Strategy, buffer_size: int = 256, source=None):
  """Load and return preprocessed and distributed datasets (main.py:18-83).

  Side effect (kept): sets ``args.train_steps`` / ``args.test_steps``."""
  size = int(getattr(args, "image_size", 256))
  resize = int(round(size * 286 / 256))
  seed = int(getattr(args, "seed", 1234))
  if source is None:
    data_dir = getattr(args, "data_dir", None)
    if data_dir:
      source = open_source(data_dir)
    else:
      sizes = None
      lim = getattr(args, "synthetic_size", None)
      if lim:
        sizes = {"trainA": lim, "trainB": lim + lim // 4,
                 "testA": max(1, lim // 8), "testB": max(1, lim // 8) + 1}
      source = SyntheticSource(raw_size=size, sizes=sizes, seed=seed)

  num_train = min(source.num_examples("trainA"), source.num_examples("trainB"))
  num_test = min(source.num_examples("testA"), source.num_examples("testB"))
  args.train_steps = ceil(num_train / args.global_batch_size)
  args.test_steps = ceil(num_test / args.global_batch_size)

  # cache() after the random map: augmentation is drawn ONCE (main.py:53-54).
  rng_a = np.random.default_rng([seed, 1])
  rng_b = np.random.default_rng([seed, 2])
  def load(split, n, rng):
    t = _materialise_native(source, split, n, rng, (resize, resize), size)
    if t is not None:
      return t
    if rng is None:
      return _materialise(source, split, n, lambda im, i: preprocess_test(im, size))
    return _materialise(source, split, n,
                        lambda im, i: preprocess_train(im, rng, (resize, resize), size))

  train_a, train_b = load("trainA", num_train, rng_a), load("trainB", num_train, rng_b)
  test_a, test_b = load("testA", num_test, None), load("testB", num_test, None)

  train_ds = PairedBatches(train_a, train_b, args.global_batch_size, strategy,
                           shuffle_buffer=buffer_size, seed=seed, prefetch=2)
  test_ds = PairedBatches(test_a, test_b, args.global_batch_size, strategy)
  # take 5 samples from the test set for plotting; NOT distributed (main.py:76-77)
  k = min(5, num_test)
  plot_ds = PairedBatches(test_a[:k], test_b[:k], 1, strategy, distribute=False)
  return train_ds, test_ds, plot_ds
