"""Reader for a TFDS-prepared ``cycle_gan/horse2zebra`` directory — the reference's only data source
(main.py:22-33: ``tfds.load("cycle_gan/horse2zebra", with_info=True, as_supervised=True)``), without
TensorFlow or tensorflow_datasets:

    <data_dir>/cycle_gan/horse2zebra/<version>/
        dataset_info.json                              split names + shard lengths
        cycle_gan-trainA.tfrecord-00000-of-00001       TFRecord shards, one tf.train.Example per image:
        cycle_gan-trainB.tfrecord-...                    features { image: bytes_list[encoded PNG/JPEG],
        cycle_gan-testA.tfrecord-... / testB ...                    label: int64_list[class id] }

TFRecord framing (uint64 length, masked crc32c of the length, payload, masked crc32c of the payload)
and the protobuf wire format are parsed with the repo's own helpers (the event-file writer and the
TensorBundle reader already use them); images are decoded with Pillow. The shards are indexed once
(offset + length per record) and read lazily. ``write_tfds_dir`` builds such a directory from arrays
(used by the tests; also handy for converting an image folder once).
"""
from __future__ import annotations

import io
import json
import os
import re
import struct
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from ..utils.summary import masked_crc32c
from ..utils.tensorbundle import _field_bytes, _field_varint, _parse

SPLITS = ("trainA", "trainB", "testA", "testB")
_SHARD = re.compile(r"^(?P<ds>[\w\-]+)-(?P<split>\w+)\.tfrecord-(?P<i>\d{5})-of-(?P<n>\d{5})$")


# ------------------------------------------------------------------ tf.train.Example
def decode_example(buf: bytes) -> Dict[str, list]:
  """``{feature name: [values]}`` of a serialized ``tf.train.Example`` (bytes / int64 / float lists)."""
  out: Dict[str, list] = {}
  for num, wt, features in _parse(buf):
    if num != 1 or wt != 2:
      continue
    for fnum, fwt, entry in _parse(features):          # map<string, Feature> entries
      if fnum != 1 or fwt != 2:
        continue
      key, feat = None, b""
      for enum, ewt, v in _parse(entry):
        if enum == 1:
          key = bytes(v).decode("utf-8")
        elif enum == 2:
          feat = bytes(v)
      vals: list = []
      for knum, kwt, lst in _parse(feat):
        if kwt != 2:
          continue
        if knum == 1:                                  # BytesList
          vals += [bytes(v) for n_, w_, v in _parse(lst) if n_ == 1]
        elif knum == 3:                                # Int64List (packed or not)
          for n_, w_, v in _parse(lst):
            if n_ != 1:
              continue
            if w_ == 0:
              vals.append(_signed(v))
            else:
              vals += _unpack_varints(bytes(v))
        elif knum == 2:                                # FloatList (packed or not)
          for n_, w_, v in _parse(lst):
            if n_ != 1:
              continue
            vals += list(struct.unpack(f"<{len(v) // 4}f", v)) if w_ == 2 else [struct.unpack("<f", struct.pack("<I", v))[0]]
      if key is not None:
        out[key] = vals
  return out


def _signed(v: int) -> int:
  return v - (1 << 64) if v >= 1 << 63 else v


def _unpack_varints(buf: bytes) -> List[int]:
  out, pos = [], 0
  while pos < len(buf):
    shift = v = 0
    while True:
      b = buf[pos]
      pos += 1
      v |= (b & 0x7F) << shift
      if not b & 0x80:
        break
      shift += 7
    out.append(_signed(v))
  return out


def encode_example(image_bytes: bytes, label: int) -> bytes:
  """Serialized ``tf.train.Example{image: bytes, label: int64}`` (what TFDS writes for cycle_gan)."""
  def entry(key: str, feature: bytes) -> bytes:
    return _field_bytes(1, _field_bytes(1, key.encode()) + _field_bytes(2, feature))
  img = _field_bytes(1, _field_bytes(1, image_bytes))                  # Feature.bytes_list.value
  lab = _field_bytes(3, _field_varint(1, label))                       # Feature.int64_list.value
  return _field_bytes(1, entry("image", img) + entry("label", lab))   # Example.features


# ------------------------------------------------------------------------ TFRecord
def index_tfrecord(path: str, verify: bool = False) -> List[Tuple[int, int]]:
  """[(payload offset, payload length)] of every record of one shard."""
  out = []
  size = os.path.getsize(path)
  with open(path, "rb") as f:
    pos = 0
    while pos < size:
      head = f.read(12)
      if len(head) < 12:
        raise ValueError(f"{path}: truncated record header at byte {pos}")
      n, = struct.unpack("<Q", head[:8])
      if verify and struct.unpack("<I", head[8:])[0] != masked_crc32c(head[:8]):
        raise ValueError(f"{path}: corrupt length checksum at byte {pos}")
      if pos + 12 + n + 4 > size:
        raise ValueError(f"{path}: truncated record at byte {pos}")
      out.append((pos + 12, n))
      if verify:
        data = f.read(n)
        if struct.unpack("<I", f.read(4))[0] != masked_crc32c(data):
          raise ValueError(f"{path}: corrupt record at byte {pos}")
      else:
        f.seek(n + 4, os.SEEK_CUR)
      pos += 12 + n + 4
  return out


def write_tfrecord(path: str, records: Iterable[bytes]) -> None:
  with open(path, "wb") as f:
    for r in records:
      head = struct.pack("<Q", len(r))
      f.write(head + struct.pack("<I", masked_crc32c(head)) + r + struct.pack("<I", masked_crc32c(r)))


# -------------------------------------------------------------------------- source
def find_tfds_dir(data_dir: str, name: str = "cycle_gan/horse2zebra") -> Optional[str]:
  """The versioned builder directory under ``data_dir`` (or ``data_dir`` itself), or None."""
  def is_builder(d: str) -> bool:
    return os.path.isfile(os.path.join(d, "dataset_info.json")) or any(_SHARD.match(f) for f in os.listdir(d))
  if not os.path.isdir(data_dir):
    return None
  if is_builder(data_dir):
    return data_dir
  base = os.path.join(data_dir, *name.split("/"))
  if os.path.isdir(base):
    if is_builder(base):
      return base
    vers = sorted((d for d in os.listdir(base) if os.path.isdir(os.path.join(base, d)) and is_builder(os.path.join(base, d))),
                  key=lambda v: [int(x) if x.isdigit() else x for x in re.split(r"[.]", v)])
    if vers:
      return os.path.join(base, vers[-1])
  return None


class TFDSSource:
  """``num_examples(split)`` / ``get(split, i) -> uint8 [H, W, 3]`` over the TFRecord shards."""

  def __init__(self, data_dir: str, name: str = "cycle_gan/horse2zebra", verify: bool = False):
    d = find_tfds_dir(data_dir, name)
    if d is None:
      raise FileNotFoundError(f"no TFDS-prepared {name} under {data_dir}")
    self.dir = d
    shards: Dict[str, List[Tuple[int, str]]] = {}
    for f in os.listdir(d):
      m = _SHARD.match(f)
      if m:
        shards.setdefault(m.group("split"), []).append((int(m.group("i")), os.path.join(d, f)))
    self.records: Dict[str, List[Tuple[str, int, int]]] = {}
    for split, lst in shards.items():
      recs = []
      for _, path in sorted(lst):
        recs += [(path, off, n) for off, n in index_tfrecord(path, verify)]
      self.records[split] = recs
    missing = [s for s in SPLITS if s not in self.records]
    if missing:
      raise FileNotFoundError(f"{d}: no shards for split(s) {missing}")
    self.info = {}
    info_path = os.path.join(d, "dataset_info.json")
    if os.path.isfile(info_path):
      self.info = json.load(open(info_path))
      for sp in self.info.get("splits", []):
        want = sum(int(x) for x in sp.get("shardLengths", []))
        have = len(self.records.get(sp.get("name"), []))
        if want and sp.get("name") in self.records and want != have:
          raise ValueError(f"{d}: split {sp['name']} has {have} records, dataset_info.json says {want}")
    self._files: Dict[str, object] = {}

  def num_examples(self, split: str) -> int:
    return len(self.records[split])

  def _read(self, split: str, index: int) -> bytes:
    path, off, n = self.records[split][index]
    f = self._files.get(path)
    if f is None:
      f = self._files[path] = open(path, "rb")
    f.seek(off)
    return f.read(n)

  def example(self, split: str, index: int) -> Dict[str, list]:
    return decode_example(self._read(split, index))

  def get(self, split: str, index: int) -> np.ndarray:
    from PIL import Image
    ex = self.example(split, index)
    return np.asarray(Image.open(io.BytesIO(ex["image"][0])).convert("RGB"))

  def label(self, split: str, index: int) -> int:
    return int(self.example(split, index).get("label", [0])[0])


def write_tfds_dir(data_dir: str, splits: Dict[str, Sequence[np.ndarray]], version: str = "2.0.0",
                   name: str = "cycle_gan/horse2zebra", shards: int = 1, fmt: str = "PNG") -> str:
  """Build ``<data_dir>/<name>/<version>/`` with one (or more) TFRecord shard(s) per split."""
  from PIL import Image
  d = os.path.join(data_dir, *name.split("/"), version)
  os.makedirs(d, exist_ok=True)
  ds = name.split("/")[0]
  info = {"name": ds, "configName": name.split("/")[-1], "version": version, "splits": []}
  for split, images in splits.items():
    label = 0 if split.endswith("A") else 1
    recs = []
    for im in images:
      buf = io.BytesIO()
      Image.fromarray(np.asarray(im, dtype=np.uint8)).save(buf, format=fmt)
      recs.append(encode_example(buf.getvalue(), label))
    per = -(-len(recs) // shards) if recs else 0
    lengths = []
    for i in range(shards):
      part = recs[i * per:(i + 1) * per]
      write_tfrecord(os.path.join(d, f"{ds}-{split}.tfrecord-{i:05d}-of-{shards:05d}"), part)
      lengths.append(str(len(part)))
    info["splits"].append({"name": split, "shardLengths": lengths, "numBytes": str(sum(len(r) for r in recs))})
  json.dump(info, open(os.path.join(d, "dataset_info.json"), "w"), indent=1)
  return d
