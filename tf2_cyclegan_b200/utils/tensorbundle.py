"""TensorBundle (TensorFlow V2 checkpoint) reader / writer, from the format description only.

The reference saves with ``tf.train.Checkpoint.write(prefix)`` (main.py:157-160), which
produces a *TensorBundle*:

``<prefix>.index``                 an SSTable (LevelDB table format): key ``""`` ->
                                   ``BundleHeaderProto``, then one ``BundleEntryProto`` per
                                   tensor key (dtype, shape, shard, offset, size, masked crc32c)
``<prefix>.data-00000-of-00001``   the raw little-endian tensor bytes, back to back

plus the key ``_CHECKPOINTABLE_OBJECT_GRAPH`` (a DT_STRING scalar holding a
``TrackableObjectGraph`` proto) that ``Checkpoint.read`` walks to match variables.

This module re-implements both containers (table blocks with prefix-compressed keys and restart
points, block trailers with masked CRC-32C, footer + magic; protobuf wire encoding; the string
tensor layout) so that checkpoints written by the reference can be loaded here and checkpoints
written here use the reference's on-disk format. TensorFlow itself is not installable in this
environment, so interchange is verified by round trips and golden byte checks of the container
primitives, not against a TensorFlow build.
"""
from __future__ import annotations

import json
import os
import struct
from typing import Any, Dict, Iterable, List, Optional, Tuple

import numpy as np
import torch

from ..ops.hostlib import crc32c, unmask_crc32c

DATA_SUFFIX = ".data-00000-of-00001"
OBJECT_GRAPH_KEY = "_CHECKPOINTABLE_OBJECT_GRAPH"
EXTRA_KEY = "_b200_cyclegan/extra"          # JSON side-car (epoch, RNG); unknown to TF, ignored by it
TABLE_MAGIC = 0xDB4775248B80FB57
BLOCK_SIZE = 262144
RESTART_INTERVAL = 16

# tensorflow/core/framework/types.proto
DT_FLOAT, DT_DOUBLE, DT_INT32, DT_UINT8, DT_INT16, DT_INT8, DT_STRING = 1, 2, 3, 4, 5, 6, 7
DT_INT64, DT_BOOL, DT_BFLOAT16, DT_HALF = 9, 10, 14, 19
_NP_TO_DT = {np.dtype("float32"): DT_FLOAT, np.dtype("float64"): DT_DOUBLE,
             np.dtype("int32"): DT_INT32, np.dtype("uint8"): DT_UINT8, np.dtype("int16"): DT_INT16,
             np.dtype("int8"): DT_INT8, np.dtype("int64"): DT_INT64, np.dtype("bool"): DT_BOOL,
             np.dtype("float16"): DT_HALF}
_DT_TO_NP = {v: k for k, v in _NP_TO_DT.items()}


def _mask(crc: int) -> int:
  return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ---------------------------------------------------------------- protobuf wire
def _varint(n: int) -> bytes:
  if n < 0:
    n += 1 << 64
  out = bytearray()
  while True:
    b = n & 0x7F
    n >>= 7
    if n:
      out.append(b | 0x80)
    else:
      out.append(b)
      return bytes(out)


def _read_varint(buf: bytes, pos: int) -> Tuple[int, int]:
  shift = result = 0
  while True:
    b = buf[pos]
    pos += 1
    result |= (b & 0x7F) << shift
    if not b & 0x80:
      return result, pos
    shift += 7


def _field_varint(num: int, value: int) -> bytes:
  return _varint(num << 3) + _varint(value)


def _field_bytes(num: int, value: bytes) -> bytes:
  return _varint((num << 3) | 2) + _varint(len(value)) + value


def _field_fixed32(num: int, value: int) -> bytes:
  return _varint((num << 3) | 5) + struct.pack("<I", value)


def _parse(buf: bytes) -> List[Tuple[int, int, Any]]:
  """[(field number, wire type, value)] of one message (no schema)."""
  out, pos = [], 0
  while pos < len(buf):
    tag, pos = _read_varint(buf, pos)
    num, wt = tag >> 3, tag & 7
    if wt == 0:
      v, pos = _read_varint(buf, pos)
    elif wt == 1:
      v = struct.unpack_from("<Q", buf, pos)[0]
      pos += 8
    elif wt == 2:
      n, pos = _read_varint(buf, pos)
      v = buf[pos:pos + n]
      pos += n
    elif wt == 5:
      v = struct.unpack_from("<I", buf, pos)[0]
      pos += 4
    else:
      raise ValueError(f"unsupported protobuf wire type {wt}")
    out.append((num, wt, v))
  return out


def _signed64(v: int) -> int:
  return v - (1 << 64) if v >= 1 << 63 else v


def encode_shape(shape: Iterable[int]) -> bytes:
  """TensorShapeProto: repeated Dim dim = 2 { int64 size = 1 }."""
  return b"".join(_field_bytes(2, _field_varint(1, int(d)) if d else b"") for d in shape)


def decode_shape(buf: bytes) -> Tuple[int, ...]:
  dims = []
  for num, _, v in _parse(buf):
    if num == 2:
      size = 0
      for n2, _, v2 in _parse(v):
        if n2 == 1:
          size = _signed64(v2)
      dims.append(size)
  return tuple(dims)


def encode_entry(dtype: int, shape: Iterable[int], offset: int, size: int, crc_masked: int) -> bytes:
  """BundleEntryProto {dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6 (fixed32)}."""
  out = _field_varint(1, dtype) + _field_bytes(2, encode_shape(shape))
  if offset:
    out += _field_varint(4, offset)
  if size:
    out += _field_varint(5, size)
  return out + _field_fixed32(6, crc_masked)


def decode_entry(buf: bytes) -> Dict[str, Any]:
  e = {"dtype": 0, "shape": (), "shard_id": 0, "offset": 0, "size": 0, "crc32c": 0, "slices": 0}
  for num, _, v in _parse(buf):
    if num == 1:
      e["dtype"] = v
    elif num == 2:
      e["shape"] = decode_shape(v)
    elif num == 3:
      e["shard_id"] = v
    elif num == 4:
      e["offset"] = v
    elif num == 5:
      e["size"] = v
    elif num == 6:
      e["crc32c"] = v
    elif num == 7:
      e["slices"] += 1
  return e


def encode_header(num_shards: int = 1) -> bytes:
  """BundleHeaderProto {num_shards=1, endianness=2 (LITTLE=0), version=3 {producer=1}}."""
  return _field_varint(1, num_shards) + _field_bytes(3, _field_varint(1, 1))


# ------------------------------------------------------------------- SSTable
class _BlockBuilder:
  def __init__(self):
    self.buf = bytearray()
    self.restarts = [0]
    self.counter = 0
    self.last_key = b""

  def add(self, key: bytes, value: bytes) -> None:
    shared = 0
    if self.counter < RESTART_INTERVAL:
      m = min(len(key), len(self.last_key))
      while shared < m and key[shared] == self.last_key[shared]:
        shared += 1
    else:
      self.restarts.append(len(self.buf))
      self.counter = 0
    self.buf += _varint(shared) + _varint(len(key) - shared) + _varint(len(value))
    self.buf += key[shared:] + value
    self.last_key = key
    self.counter += 1

  def size_estimate(self) -> int:
    return len(self.buf) + 4 * len(self.restarts) + 4

  def empty(self) -> bool:
    return not self.buf

  def finish(self) -> bytes:
    return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + \
        struct.pack("<I", len(self.restarts))


def _block_handle(offset: int, size: int) -> bytes:
  return _varint(offset) + _varint(size)


def write_table(path: str, items: List[Tuple[bytes, bytes]]) -> None:
  """Write sorted (key, value) pairs as an uncompressed LevelDB-format table (what
  ``BundleWriter::Finish`` produces with ``table::kNoCompression``)."""
  keys = [k for k, _ in items]
  if keys != sorted(keys) or len(set(keys)) != len(keys):
    raise ValueError("table keys must be unique and sorted")
  out = bytearray()

  def emit(block: bytes) -> Tuple[int, int]:
    off = len(out)
    trailer_type = b"\0"                                   # kNoCompression
    crc = crc32c(trailer_type, crc32c(block))
    out.extend(block + trailer_type + struct.pack("<I", _mask(crc)))
    return off, len(block)

  index = _BlockBuilder()
  data = _BlockBuilder()
  for key, value in items:
    data.add(key, value)
    if data.size_estimate() >= BLOCK_SIZE:
      off, size = emit(data.finish())
      index.add(data.last_key, _block_handle(off, size))   # separator = last key of the block
      data = _BlockBuilder()
  if not data.empty():
    off, size = emit(data.finish())
    index.add(data.last_key, _block_handle(off, size))
  meta_off, meta_size = emit(_BlockBuilder().finish())       # empty metaindex block
  idx_off, idx_size = emit(index.finish())
  footer = _block_handle(meta_off, meta_size) + _block_handle(idx_off, idx_size)
  footer += b"\0" * (40 - len(footer))
  footer += struct.pack("<II", TABLE_MAGIC & 0xFFFFFFFF, TABLE_MAGIC >> 32)
  out.extend(footer)
  with open(path, "wb") as f:
    f.write(out)
    f.flush()
    os.fsync(f.fileno())


def _snappy_decompress(buf: bytes) -> bytes:
  n, pos = _read_varint(buf, 0)
  out = bytearray()
  while pos < len(buf):
    tag = buf[pos]
    pos += 1
    kind = tag & 3
    if kind == 0:                                           # literal
      ln = tag >> 2
      if ln >= 60:
        nb = ln - 59
        ln = int.from_bytes(buf[pos:pos + nb], "little")
        pos += nb
      ln += 1
      out += buf[pos:pos + ln]
      pos += ln
      continue
    if kind == 1:
      ln = ((tag >> 2) & 7) + 4
      off = ((tag >> 5) << 8) | buf[pos]
      pos += 1
    elif kind == 2:
      ln = (tag >> 2) + 1
      off = int.from_bytes(buf[pos:pos + 2], "little")
      pos += 2
    else:
      ln = (tag >> 2) + 1
      off = int.from_bytes(buf[pos:pos + 4], "little")
      pos += 4
    if off == 0 or off > len(out):
      raise ValueError("corrupt snappy stream")
    for _ in range(ln):                                     # may overlap (run-length)
      out.append(out[-off])
  if len(out) != n:
    raise ValueError("corrupt snappy stream (length)")
  return bytes(out)


def _read_block(blob: bytes, offset: int, size: int, verify: bool = True) -> bytes:
  block = blob[offset:offset + size]
  ctype = blob[offset + size]
  stored = struct.unpack_from("<I", blob, offset + size + 1)[0]
  if verify and unmask_crc32c(stored) != crc32c(blob[offset + size:offset + size + 1], crc32c(block)):
    raise IOError("table block checksum mismatch")
  if ctype == 0:
    return block
  if ctype == 1:
    return _snappy_decompress(block)
  raise ValueError(f"unknown block compression {ctype}")


def _block_entries(block: bytes) -> List[Tuple[bytes, bytes]]:
  num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
  end = len(block) - 4 - 4 * num_restarts
  out, pos, key = [], 0, b""
  while pos < end:
    shared, pos = _read_varint(block, pos)
    non_shared, pos = _read_varint(block, pos)
    vlen, pos = _read_varint(block, pos)
    key = key[:shared] + block[pos:pos + non_shared]
    pos += non_shared
    out.append((key, block[pos:pos + vlen]))
    pos += vlen
  return out


def read_table(path: str) -> List[Tuple[bytes, bytes]]:
  with open(path, "rb") as f:
    blob = f.read()
  if len(blob) < 48:
    raise ValueError(f"{path}: too short for a table")
  lo, hi = struct.unpack_from("<II", blob, len(blob) - 8)
  if (hi << 32 | lo) != TABLE_MAGIC:
    raise ValueError(f"{path}: not an SSTable (bad magic)")
  footer = blob[len(blob) - 48:]
  pos = 0
  _, pos = _read_varint(footer, pos)
  _, pos = _read_varint(footer, pos)
  idx_off, pos = _read_varint(footer, pos)
  idx_size, pos = _read_varint(footer, pos)
  out = []
  for _, handle in _block_entries(_read_block(blob, idx_off, idx_size)):
    off, p = _read_varint(handle, 0)
    size, _ = _read_varint(handle, p)
    out.extend(_block_entries(_read_block(blob, off, size)))
  return out


# ------------------------------------------------------------ object graph
def encode_object_graph(keys: Iterable[str]) -> bytes:
  """A ``TrackableObjectGraph`` covering ``keys`` (object paths joined by '/', the part before
  ``/.ATTRIBUTES/``): nodes for every path prefix, ``children`` edges named by path component,
  one ``SerializedTensor`` attribute per variable, and ``slot_variables`` on each optimizer for
  keys of the form ``<var path>/.OPTIMIZER_SLOT/<optimizer>/<slot>``."""
  nodes: List[Dict[str, Any]] = [{"children": {}, "attrs": [], "slots": []}]

  def node_for(path: Tuple[str, ...]) -> int:
    cur = 0
    for comp in path:
      nxt = nodes[cur]["children"].get(comp)
      if nxt is None:
        nodes.append({"children": {}, "attrs": [], "slots": []})
        nxt = len(nodes) - 1
        nodes[cur]["children"][comp] = nxt
      cur = nxt
    return cur

  plain, slots = [], []
  for key in sorted(keys):
    if "/.ATTRIBUTES/" not in key:
      continue
    path, attr = key.split("/.ATTRIBUTES/")
    (slots if "/.OPTIMIZER_SLOT/" in path else plain).append((path, attr, key))
  for path, attr, key in plain:
    n = node_for(tuple(path.split("/")))
    nodes[n]["attrs"].append((attr, path.split("/")[-1], key))
  for path, attr, key in slots:
    var_path, rest = path.split("/.OPTIMIZER_SLOT/")
    opt_name, slot_name = rest.split("/")
    var_node = node_for(tuple(var_path.split("/")))
    opt_node = node_for((opt_name,))
    nodes.append({"children": {}, "attrs": [(attr, f"{var_path.split('/')[-1]}/{slot_name}", key)],
                  "slots": []})
    nodes[opt_node]["slots"].append((var_node, slot_name, len(nodes) - 1))

  out = b""
  for nd in nodes:
    body = b""
    for name, child in nd["children"].items():
      ref = (_field_varint(1, child) if child else b"") + _field_bytes(2, name.encode())
      body += _field_bytes(1, ref)
    for attr, full_name, key in nd["attrs"]:
      body += _field_bytes(2, _field_bytes(1, attr.encode()) + _field_bytes(2, full_name.encode()) +
                           _field_bytes(3, key.encode()))
    for var_node, slot_name, slot_node in nd["slots"]:
      body += _field_bytes(3, _field_varint(1, var_node) + _field_bytes(2, slot_name.encode()) +
                           _field_varint(3, slot_node))
    out += _field_bytes(1, body)
  return out


def decode_object_graph(buf: bytes) -> List[Dict[str, Any]]:
  nodes = []
  for num, _, v in _parse(buf):
    if num != 1:
      continue
    nd = {"children": {}, "attrs": [], "slots": []}
    for n2, _, v2 in _parse(v):
      f = {a: c for a, _, c in _parse(v2)}
      if n2 == 1:
        nd["children"][f.get(2, b"").decode()] = f.get(1, 0)
      elif n2 == 2:
        nd["attrs"].append((f.get(1, b"").decode(), f.get(2, b"").decode(), f.get(3, b"").decode()))
      elif n2 == 3:
        nd["slots"].append((f.get(1, 0), f.get(2, b"").decode(), f.get(3, 0)))
    nodes.append(nd)
  return nodes


# ------------------------------------------------------------------- bundle
def _string_tensor_bytes(strings: List[bytes]) -> Tuple[bytes, int]:
  """[varint64 len]* [masked crc32c of the lengths (as uint32 LE)] [bytes]*; returns the payload
  and the entry checksum (lengths, the 4 checksum bytes, then all string bytes)."""
  lengths = b"".join(_varint(len(s)) for s in strings)
  crc = 0
  for s in strings:
    crc = crc32c(struct.pack("<I", len(s)), crc)
  cks = struct.pack("<I", _mask(crc))
  crc = crc32c(cks, crc)
  for s in strings:
    crc = crc32c(s, crc)
  return lengths + cks + b"".join(strings), crc


def _as_numpy(t) -> Tuple[np.ndarray, int]:
  if isinstance(t, torch.Tensor):
    t = t.detach().cpu().contiguous()
    if t.dtype == torch.bfloat16:
      return t.view(torch.int16).numpy().reshape(tuple(t.shape)), DT_BFLOAT16
    t = t.numpy()
  arr = np.asarray(t)
  arr = np.ascontiguousarray(arr).reshape(arr.shape)        # ascontiguousarray promotes 0-d to 1-d
  if arr.dtype not in _NP_TO_DT:
    raise TypeError(f"dtype {arr.dtype} has no TensorBundle mapping")
  return arr, _NP_TO_DT[arr.dtype]


def write_bundle(prefix: str, tensors: Dict[str, Any], extra: Optional[Dict[str, Any]] = None,
                 object_graph: bool = True) -> None:
  """Write ``tensors`` (key -> torch / numpy array, or ``bytes`` for a DT_STRING scalar) as a
  one-shard TensorBundle. Both files are written under temporary names, fsynced and renamed (data
  first, index last). The TF format has no field that could tie the two files together, so a crash
  between the two renames leaves a NEW data file next to an OLD index: :func:`read_bundle` detects
  that pair (data-file length vs the index's extents, then the per-tensor crc32c) and refuses it
  instead of loading garbage. The native container (utils/checkpoint.py) carries a nonce and is
  atomic in the strict sense."""
  os.makedirs(os.path.dirname(prefix) or ".", exist_ok=True)
  payload: Dict[str, Any] = dict(tensors)
  if object_graph and OBJECT_GRAPH_KEY not in payload:
    payload[OBJECT_GRAPH_KEY] = encode_object_graph(payload.keys())
  if extra:
    payload[EXTRA_KEY] = json.dumps(extra).encode()
  items: List[Tuple[bytes, bytes]] = [(b"", encode_header(1))]
  offset = 0
  data_tmp, index_tmp = prefix + DATA_SUFFIX + ".tmp", prefix + ".index.tmp"
  with open(data_tmp, "wb") as f:
    for key in sorted(payload, key=lambda k: k.encode()):
      value = payload[key]
      if isinstance(value, (bytes, bytearray)):
        raw, crc = _string_tensor_bytes([bytes(value)])
        dtype, shape = DT_STRING, ()
      else:
        arr, dtype = _as_numpy(value)
        raw, crc, shape = arr.tobytes(), crc32c(arr), arr.shape
      items.append((key.encode(), encode_entry(dtype, shape, offset, len(raw), _mask(crc))))
      f.write(raw)
      offset += len(raw)
    f.flush()
    os.fsync(f.fileno())
  write_table(index_tmp, items)
  os.replace(data_tmp, prefix + DATA_SUFFIX)
  os.replace(index_tmp, prefix + ".index")


def is_bundle(prefix: str) -> bool:
  """True when ``<prefix>.index`` ends with the table magic number."""
  try:
    with open(prefix + ".index", "rb") as f:
      f.seek(-8, os.SEEK_END)
      lo, hi = struct.unpack("<II", f.read(8))
    return (hi << 32 | lo) == TABLE_MAGIC
  except (OSError, struct.error):
    return False


def read_bundle(prefix: str, verify: bool = True) -> Tuple[Dict[str, torch.Tensor], Dict[str, Any]]:
  """-> (tensors, info). ``info`` = {"extra": dict, "object_graph": bytes | None,
  "strings": {key: [bytes]}}. String tensors other than the side-car end up in ``strings``."""
  entries = read_table(prefix + ".index")
  if not entries or entries[0][0] != b"":
    raise ValueError(f"{prefix}.index: missing bundle header entry")
  header = {n: v for n, _, v in _parse(entries[0][1])}
  num_shards = header.get(1, 0)
  if header.get(2, 0) != 0:
    raise ValueError("big-endian bundles are not supported")
  shards: Dict[int, np.memmap] = {}

  def shard(i: int):
    if i not in shards:
      path = f"{prefix}.data-{i:05d}-of-{num_shards:05d}"
      shards[i] = np.memmap(path, dtype=np.uint8, mode="r") if os.path.getsize(path) else \
          np.zeros(0, np.uint8)
    return shards[i]

  tensors: Dict[str, torch.Tensor] = {}
  info: Dict[str, Any] = {"extra": {}, "object_graph": None, "strings": {}}
  if verify:   # a writer lays tensors out back to back: the index's extents must cover the shard exactly
    extent: Dict[int, int] = {}
    for k, v in entries[1:]:
      e = decode_entry(v)
      extent[e["shard_id"]] = max(extent.get(e["shard_id"], 0), e["offset"] + e["size"])
    for i, end in extent.items():
      have = os.path.getsize(f"{prefix}.data-{i:05d}-of-{num_shards:05d}")
      if have != end:
        raise IOError(f"{prefix}: data shard {i} is {have} bytes but the index describes {end}: index and data "
                      "files are from different writes (interrupted save?)")
  for k, v in entries[1:]:
    key = k.decode()
    e = decode_entry(v)
    if e["slices"]:
      raise NotImplementedError(f"{key}: partitioned (sliced) variables are not supported")
    raw = shard(e["shard_id"])[e["offset"]:e["offset"] + e["size"]]
    if e["dtype"] == DT_STRING:
      rawb = np.asarray(raw).tobytes()
      strings = _decode_string_tensor(rawb, int(np.prod(e["shape"], dtype=np.int64)))
      if verify and (_string_tensor_bytes(strings) != (rawb, unmask_crc32c(e["crc32c"]))):
        raise IOError(f"{prefix}: checksum mismatch for {key}")
      if key == OBJECT_GRAPH_KEY:
        info["object_graph"] = strings[0]
      elif key == EXTRA_KEY:
        info["extra"] = json.loads(strings[0].decode())
      else:
        info["strings"][key] = strings
      continue
    if verify and unmask_crc32c(e["crc32c"]) != crc32c(np.ascontiguousarray(raw)):
      raise IOError(f"{prefix}: checksum mismatch for {key}")
    if e["dtype"] == DT_BFLOAT16:
      t = torch.from_numpy(np.frombuffer(np.asarray(raw).tobytes(), dtype=np.int16).copy()).view(torch.bfloat16)
    elif e["dtype"] in _DT_TO_NP:
      t = torch.from_numpy(np.frombuffer(np.asarray(raw).tobytes(), dtype=_DT_TO_NP[e["dtype"]]).copy())
    else:
      raise TypeError(f"{key}: unsupported DataType {e['dtype']}")
    tensors[key] = torch.reshape(t, tuple(e["shape"]))
  return tensors, info


def _decode_string_tensor(raw: bytes, count: int) -> List[bytes]:
  pos, lens = 0, []
  for _ in range(count):
    n, pos = _read_varint(raw, pos)
    lens.append(n)
  pos += 4                                                   # checksum of the lengths
  out = []
  for n in lens:
    out.append(raw[pos:pos + n])
    pos += n
  return out


# --------------------------------------------- reference <-> framework state
_OPT_HYPER = ("learning_rate", "beta_1", "beta_2", "decay")


def to_reference_dtypes(state: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
  """Keras stores ``<opt>/iter`` as int64 and the hyper-parameters as float32; this framework's
  state dict keeps them as float64 scalars."""
  out = {}
  for k, v in state.items():
    head = k.split("/.ATTRIBUTES/")[0]
    if head.endswith("_optimizer/iter"):
      v = v.to(torch.int64)
    elif any(head.endswith("_optimizer/" + h) for h in _OPT_HYPER):
      v = v.to(torch.float32)
    out[k] = v
  return out


def from_reference_dtypes(state: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
  out = {}
  for k, v in state.items():
    head = k.split("/.ATTRIBUTES/")[0]
    if head.endswith("_optimizer/iter") or any(head.endswith("_optimizer/" + h) for h in _OPT_HYPER):
      v = v.to(torch.float64)
    out[k] = v
  return out
