"""TensorBoard summaries without TensorFlow: a minimal native event-file writer
plus the reference's ``Summary`` helper (cyclegan/utils.py:14-98).

Event files are TFRecord streams of hand-encoded ``tensorflow.Event`` protos
in the TF2 style (tensor-valued summaries with ``scalars`` / ``images`` plugin
metadata), so TensorBoard shows them exactly like the reference's runs:
writer 0 = ``<output_dir>/`` (training=True), writer 1 = ``<output_dir>/test``.
matplotlib is not available; the 1×3 cycle figure is composed with Pillow
(utils/plot.py).
"""
from __future__ import annotations

import io
import os
import socket
import struct
import time
from typing import List, Sequence

import numpy as np

# ------------------------------------------------------------------ crc32c
from ..ops.hostlib import crc32c, masked_crc32c  # noqa: E402  (native SSE4.2, Python fallback)


# ------------------------------------------------------- protobuf encoding
def _varint(n: int) -> bytes:
  out = bytearray()
  if n < 0:
    n += 1 << 64
  while True:
    b = n & 0x7F
    n >>= 7
    if n:
      out.append(b | 0x80)
    else:
      out.append(b)
      return bytes(out)


def _key(field: int, wire: int) -> bytes:
  return _varint((field << 3) | wire)


def _bytes_field(field: int, data: bytes) -> bytes:
  return _key(field, 2) + _varint(len(data)) + data


def _str_field(field: int, s: str) -> bytes:
  return _bytes_field(field, s.encode("utf-8"))


def _varint_field(field: int, n: int) -> bytes:
  return _key(field, 0) + _varint(n)


def _double_field(field: int, v: float) -> bytes:
  return _key(field, 1) + struct.pack("<d", v)


DT_FLOAT, DT_STRING = 1, 7
DATA_CLASS_SCALAR, DATA_CLASS_BLOB_SEQUENCE = 1, 3


def _metadata(plugin: str, data_class: int, content: bytes = b"") -> bytes:
  plugin_data = _str_field(1, plugin)
  if content:
    plugin_data += _bytes_field(2, content)
  return _bytes_field(1, plugin_data) + _varint_field(4, data_class)


def _scalar_value(tag: str, value: float) -> bytes:
  tensor = (_varint_field(1, DT_FLOAT) + _bytes_field(2, b"") +
            _bytes_field(5, struct.pack("<f", float(value))))  # packed float_val
  return (_str_field(1, tag) + _bytes_field(8, tensor) +
          _bytes_field(9, _metadata("scalars", DATA_CLASS_SCALAR)))


def _image_value(tag: str, width: int, height: int, pngs: Sequence[bytes]) -> bytes:
  n = 2 + len(pngs)
  shape = _bytes_field(2, _varint_field(1, n))  # TensorShapeProto.dim{size}
  tensor = _varint_field(1, DT_STRING) + _bytes_field(2, shape)
  for s in (str(width).encode(), str(height).encode(), *pngs):
    tensor += _bytes_field(8, s)
  # ImagePluginData{max_images_per_step = 1: len(pngs)}
  content = _varint_field(1, len(pngs))
  return (_str_field(1, tag) + _bytes_field(8, tensor) +
          _bytes_field(9, _metadata("images", DATA_CLASS_BLOB_SEQUENCE, content)))


def _event(step: int, summary_value: bytes = b"", file_version: str = "") -> bytes:
  ev = _double_field(1, time.time()) + _varint_field(2, step)
  if file_version:
    ev += _str_field(3, file_version)
  if summary_value:
    ev += _bytes_field(5, _bytes_field(1, summary_value))
  return ev


class EventFileWriter:
  """Append-only ``events.out.tfevents.*`` writer (TFRecord framing)."""
  _uid = 0

  def __init__(self, logdir: str):
    os.makedirs(logdir, exist_ok=True)
    EventFileWriter._uid += 1
    fname = "events.out.tfevents.%010d.%s.%d.%d.v2" % (
        int(time.time()), socket.gethostname(), os.getpid(), EventFileWriter._uid)
    self.path = os.path.join(logdir, fname)
    self._f = open(self.path, "ab")
    self._write(_event(0, file_version="brain.Event:2"))
    self.flush()

  def _write(self, record: bytes) -> None:
    header = struct.pack("<Q", len(record))
    self._f.write(header + struct.pack("<I", masked_crc32c(header)) + record +
                  struct.pack("<I", masked_crc32c(record)))

  def scalar(self, tag: str, value: float, step: int) -> None:
    self._write(_event(step, _scalar_value(tag, value)))

  def images(self, tag: str, width: int, height: int, pngs: Sequence[bytes], step: int) -> None:
    self._write(_event(step, _image_value(tag, width, height, pngs)))

  def flush(self) -> None:
    self._f.flush()

  def close(self) -> None:
    if not self._f.closed:
      self._f.flush()
      self._f.close()


def encode_png(image: np.ndarray) -> bytes:
  from PIL import Image
  buf = io.BytesIO()
  Image.fromarray(image).save(buf, format="PNG")
  return buf.getvalue()


class Summary:
  """Helper class to write TensorBoard summaries (cyclegan/utils.py:14-98).

  Only the chief replica should construct it with ``enabled=True``; other
  ranks get a no-op object so call sites stay unconditional.
  """

  def __init__(self, output_dir: str, enabled: bool = True):
    self.dpi = 120
    self.enabled = enabled
    self.writers: List[EventFileWriter] = []
    if enabled:
      self.writers = [EventFileWriter(output_dir),
                      EventFileWriter(os.path.join(output_dir, "test"))]

  def get_writer(self, training: bool) -> EventFileWriter:
    return self.writers[0 if training else 1]

  def scalar(self, tag, value, step: int = 0, training: bool = False):
    if not self.enabled:
      return
    self.get_writer(training).scalar(tag, float(value), step)
    self.get_writer(training).flush()

  def image(self, tag, values, step: int = 0, training: bool = False):
    """values: uint8 array [N, H, W, C] (C ∈ {1, 3, 4}); all N are written."""
    if not self.enabled:
      return
    values = np.asarray(values)
    pngs = [encode_png(v if v.shape[-1] != 1 else v[..., 0]) for v in values]
    self.get_writer(training).images(tag, values.shape[2], values.shape[1], pngs, step)
    self.get_writer(training).flush()

  def figure(self, tag, figure, step: int = 0, training: bool = False, close: bool = True):
    """Write a rendered figure (PIL.Image or uint8 HWC array) as an image summary.

    The reference rasterises a matplotlib figure to RGBA PNG first
    (utils.py:39-59); ours already is a raster."""
    arr = np.asarray(figure.convert("RGBA") if hasattr(figure, "convert") else figure)
    self.image(tag, arr[None, ...], step=step, training=training)

  def image_cycle(self, tag: str, images: List[np.ndarray], labels: List[str],
                  step: int = 0, training: bool = False):
    """Plot image cycle to TensorBoard: one 1×3 panel per sample (utils.py:61-98)."""
    assert len(images) == len(labels) == 3
    from .plot import cycle_figure
    for sample in range(len(images[0])):
      fig = cycle_figure([im[sample] for im in images], labels, dpi=self.dpi)
      self.figure(tag=f"{tag}/sample_#{sample:03d}", figure=fig, step=step,
                  training=training, close=True)

  def close(self):
    for w in self.writers:
      w.close()
