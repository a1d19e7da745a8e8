"""TFDS-prepared cycle_gan/horse2zebra directory -> datasets, without TensorFlow (main.py:22-33)."""
import types

import numpy as np
import pytest
import torch

from tf2_cyclegan_b200.data import tfds
from tf2_cyclegan_b200.data.pipeline import FolderSource, get_datasets, open_source
from tf2_cyclegan_b200.parallel.strategy import Strategy


def _images(n, size, seed):
  rng = np.random.default_rng(seed)
  return [rng.integers(0, 256, (size, size, 3), dtype=np.uint8) for _ in range(n)]


def test_example_roundtrip_and_packed_int64():
  ex = tfds.decode_example(tfds.encode_example(b"\x89PNG-bytes", 1))
  assert ex == {"image": [b"\x89PNG-bytes"], "label": [1]}
  # packed int64 list (what the C++ protobuf writer emits): field 1, wire type 2
  from tf2_cyclegan_b200.utils.tensorbundle import _field_bytes, _varint
  packed = _field_bytes(3, _field_bytes(1, _varint(7) + _varint(300)))
  entry = _field_bytes(1, _field_bytes(1, b"label") + _field_bytes(2, packed))
  assert tfds.decode_example(_field_bytes(1, entry)) == {"label": [7, 300]}


def test_tfds_dir_to_datasets(tmp_path):
  splits = {"trainA": _images(5, 16, 1), "trainB": _images(7, 16, 2), "testA": _images(3, 16, 3),
            "testB": _images(2, 16, 4)}
  d = tfds.write_tfds_dir(str(tmp_path), splits, shards=2)
  assert tfds.find_tfds_dir(str(tmp_path)) == d
  src = open_source(str(tmp_path))
  assert isinstance(src, tfds.TFDSSource)
  assert [src.num_examples(s) for s in tfds.SPLITS] == [5, 7, 3, 2]
  for split, imgs in splits.items():
    for i, im in enumerate(imgs):                       # PNG is lossless: exact pixels, shard order kept
      assert np.array_equal(src.get(split, i), im)
    assert src.label(split, 0) == (0 if split.endswith("A") else 1)
  # split sizes -> steps exactly as main.py:29-33 (min over the two domains, ceil by the global batch)
  args = types.SimpleNamespace(data_dir=str(tmp_path), image_size=16, global_batch_size=2, seed=1234)
  train_ds, test_ds, plot_ds = get_datasets(args, Strategy(device="cpu"))
  assert args.train_steps == 3 and args.test_steps == 1
  x, y = next(iter(test_ds))
  assert x.shape == (2, 16, 16, 3) and y.shape == (2, 16, 16, 3)
  want = torch.from_numpy(np.stack(splits["testA"][:2])).float() / 127.5 - 1
  assert torch.allclose(x.cpu(), want, atol=1e-5)      # test map: resize (identity here) + normalise
  assert len(list(iter(train_ds))) == 3


def test_tfds_detects_corruption_and_count_mismatch(tmp_path):
  splits = {s: _images(2, 8, i) for i, s in enumerate(tfds.SPLITS)}
  d = tfds.write_tfds_dir(str(tmp_path), splits)
  shard = f"{d}/cycle_gan-trainA.tfrecord-00000-of-00001"
  blob = bytearray(open(shard, "rb").read())
  blob[40] ^= 0xFF
  open(shard, "wb").write(bytes(blob))
  with pytest.raises(ValueError):
    tfds.TFDSSource(str(tmp_path), verify=True)
  tfds.write_tfds_dir(str(tmp_path), splits)           # restore, then drop a record from the info
  import json
  info = json.load(open(f"{d}/dataset_info.json"))
  info["splits"][0]["shardLengths"] = ["3"]
  json.dump(info, open(f"{d}/dataset_info.json", "w"))
  with pytest.raises(ValueError):
    tfds.TFDSSource(str(tmp_path))


def test_folder_layout_still_wins(tmp_path):
  from PIL import Image
  for s in tfds.SPLITS:
    (tmp_path / s).mkdir()
    Image.fromarray(_images(1, 8, 0)[0]).save(tmp_path / s / "a.png")
  assert isinstance(open_source(str(tmp_path)), FolderSource)
