"""TensorBundle / SSTable container (utils/tensorbundle.py): golden bytes of the primitives whose
encodings are fixed by the formats (protobuf wire, LevelDB block / footer, masked crc32c), round
trips, corruption detection, and the CycleGAN facade saving / resuming in the reference's format."""
import os
import struct

import pytest
import torch

from tf2_cyclegan_b200.utils import checkpoint as ckpt
from tf2_cyclegan_b200.utils import tensorbundle as tb


def test_wire_primitives_golden():
  assert tb._varint(0) == b"\x00" and tb._varint(300) == b"\xac\x02"
  assert tb._read_varint(b"\xac\x02", 0) == (300, 2)
  # BundleHeaderProto{num_shards: 1, version{producer: 1}}
  assert tb.encode_header(1) == bytes.fromhex("08011a020801")
  # TensorShapeProto [7, 7, 3, 64]
  assert tb.encode_shape((7, 7, 3, 64)) == bytes.fromhex("12020807" "12020807" "12020803" "12020840")
  e = tb.encode_entry(tb.DT_FLOAT, (3, 4), offset=128, size=48, crc_masked=0xDEADBEEF)
  assert e == bytes.fromhex("0801" "1208" "12020803" "12020804" "208001" "2830" "35efbeadde")
  d = tb.decode_entry(e)
  assert (d["dtype"], d["shape"], d["offset"], d["size"], d["crc32c"]) == (1, (3, 4), 128, 48, 0xDEADBEEF)
  # scalar entry at offset 0: shape present but empty, zero fields omitted
  assert tb.encode_entry(tb.DT_INT64, (), 0, 8, 1) == bytes.fromhex("0809" "1200" "2808" "3501000000")


def test_table_layout_and_roundtrip(tmp_path):
  items = [(b"", b"hdr")] + [(f"key/{i:04d}".encode(), os.urandom(i % 50)) for i in range(100)]
  path = str(tmp_path / "t.index")
  tb.write_table(path, items)
  blob = open(path, "rb").read()
  assert blob[-8:] == bytes.fromhex("57fb808b247547db")          # kTableMagicNumber, little endian
  assert len(blob) >= 48
  # first entry of the first data block: shared=0, key_len=0, value_len=3, "hdr"
  assert blob[:6] == b"\x00\x00\x03hdr"
  # second entry: restart interval 16 -> prefix compression against the (empty) previous key
  assert blob[6:9] == bytes([0, 8, 0]) and blob[9:17] == b"key/0000"
  assert tb.read_table(path) == items
  # > one data block
  big = [(f"{i:06d}".encode(), bytes(3000)) for i in range(200)]
  tb.write_table(path, big)
  assert tb.read_table(path) == big
  with pytest.raises(ValueError):
    tb.write_table(path, [(b"b", b""), (b"a", b"")])


def test_table_checksum_detects_corruption(tmp_path):
  path = str(tmp_path / "t.index")
  tb.write_table(path, [(b"", b"h"), (b"k", b"value")])
  blob = bytearray(open(path, "rb").read())
  blob[3] ^= 0x40
  open(path, "wb").write(blob)
  with pytest.raises(IOError):
    tb.read_table(path)


def test_snappy_decoder():
  # literal "abcd" + copy(offset 4, len 8) -> "abcdabcdabcd"
  stream = bytes([12, (4 - 1) << 2]) + b"abcd" + bytes([((8 - 4) << 2) | 1, 4])
  assert tb._snappy_decompress(stream) == b"abcdabcdabcd"


def test_string_tensor_layout():
  raw, crc = tb._string_tensor_bytes([b"hello"])
  assert raw[0] == 5 and raw[5:] == b"hello" and len(raw) == 1 + 4 + 5
  lens_crc = tb.crc32c(struct.pack("<I", 5))
  assert struct.unpack("<I", raw[1:5])[0] == tb._mask(lens_crc)
  assert crc == tb.crc32c(b"hello", tb.crc32c(raw[1:5], lens_crc))
  assert tb._decode_string_tensor(raw, 1) == [b"hello"]


def _state():
  g = torch.Generator().manual_seed(0)
  sd = {}
  for m, opt in (("G", "G_optimizer"), ("X", "X_optimizer")):
    for i, shape in enumerate([(7, 7, 3, 64), (64,), (64,)]):
      name = ["kernel", "gamma", "beta"][i]
      k = f"{m}/layer_with_weights-{min(i, 1)}/{name}"
      sd[f"{k}/.ATTRIBUTES/VARIABLE_VALUE"] = torch.randn(shape, generator=g)
      for slot in "mv":
        sd[f"{k}/.OPTIMIZER_SLOT/{opt}/{slot}/.ATTRIBUTES/VARIABLE_VALUE"] = torch.randn(shape, generator=g)
    sd[f"{opt}/iter/.ATTRIBUTES/VARIABLE_VALUE"] = torch.tensor(17.0, dtype=torch.float64)
    for h, v in (("learning_rate", 2e-4), ("beta_1", 0.5), ("beta_2", 0.9), ("decay", 0.0)):
      sd[f"{opt}/{h}/.ATTRIBUTES/VARIABLE_VALUE"] = torch.tensor(v, dtype=torch.float64)
  return sd


def test_bundle_roundtrip_and_reference_dtypes(tmp_path):
  prefix = str(tmp_path / "checkpoints" / "checkpoint")
  sd = _state()
  ckpt.save(prefix, sd, extra={"epoch": 7}, fmt="tensorbundle")
  assert sorted(os.listdir(os.path.dirname(prefix))) == ["checkpoint.data-00000-of-00001", "checkpoint.index"]
  assert tb.is_bundle(prefix) and ckpt.exists(prefix)
  # on disk: Keras dtypes (iter int64, hyper-parameters float32); data file = tensors back to back
  raw, info = tb.read_bundle(prefix)
  assert raw["G_optimizer/iter/.ATTRIBUTES/VARIABLE_VALUE"].dtype == torch.int64
  assert raw["G_optimizer/beta_1/.ATTRIBUTES/VARIABLE_VALUE"].dtype == torch.float32
  entries = dict(tb.read_table(prefix + ".index"))
  total = sum(tb.decode_entry(v)["size"] for k, v in entries.items() if k)
  assert os.path.getsize(prefix + tb.DATA_SUFFIX) == total
  # through the auto-detecting loader: framework dtypes, identical values, extras preserved
  back, extra = ckpt.load(prefix)
  assert extra == {"epoch": 7}
  assert set(back) == set(sd)
  for k in sd:
    assert back[k].dtype == sd[k].dtype and back[k].shape == sd[k].shape
    assert torch.allclose(back[k].double(), sd[k].double(), rtol=1e-7, atol=0), k


def test_object_graph_structure(tmp_path):
  sd = _state()
  nodes = tb.decode_object_graph(tb.encode_object_graph(sd.keys()))
  root = nodes[0]
  assert set(root["children"]) == {"G", "X", "G_optimizer", "X_optimizer"}
  layer = nodes[nodes[root["children"]["G"]]["children"]["layer_with_weights-0"]]
  kernel = nodes[layer["children"]["kernel"]]
  assert kernel["attrs"] == [("VARIABLE_VALUE", "kernel",
                              "G/layer_with_weights-0/kernel/.ATTRIBUTES/VARIABLE_VALUE")]
  opt = nodes[root["children"]["G_optimizer"]]
  assert set(opt["children"]) == {"iter", "learning_rate", "beta_1", "beta_2", "decay"}
  assert len(opt["slots"]) == 6                        # 3 variables x {m, v}
  var_node, slot_name, slot_node = opt["slots"][0]
  assert slot_name in ("m", "v") and nodes[slot_node]["attrs"][0][2].count("/.OPTIMIZER_SLOT/G_optimizer/") == 1
  # every key is reachable exactly once
  keys = [a[2] for nd in nodes for a in nd["attrs"]]
  assert sorted(keys) == sorted(sd.keys())


def test_bundle_detects_data_corruption(tmp_path):
  prefix = str(tmp_path / "c")
  tb.write_bundle(prefix, {"a/.ATTRIBUTES/VARIABLE_VALUE": torch.arange(100, dtype=torch.float32)})
  with open(prefix + tb.DATA_SUFFIX, "r+b") as f:
    f.seek(40)
    f.write(b"\xff")
  with pytest.raises(IOError):
    tb.read_bundle(prefix)


def test_facade_saves_and_resumes_in_reference_format(tmp_path):
  from tf2_cyclegan_b200.cli import parse_args
  from tf2_cyclegan_b200.engine.cyclegan import CycleGAN
  from tf2_cyclegan_b200.parallel.strategy import Strategy
  argv = ["--device", "cpu", "--backend", "oracle", "--dtype", "fp32", "--image_size", "32",
          "--num_residual_blocks", "1", "--output_dir", str(tmp_path), "--checkpoint_format", "tensorbundle"]
  args = parse_args(argv)
  strategy = Strategy(device="cpu")
  args.global_batch_size = 1
  torch.manual_seed(0)
  gan = CycleGAN(args, strategy)
  x, y = torch.rand(1, 32, 32, 3) * 2 - 1, torch.rand(1, 32, 32, 3) * 2 - 1
  gan.distributed_train_step(x, y)
  gan.save_checkpoint(extra={"epoch": 3})
  assert tb.is_bundle(gan.checkpoint_prefix)
  raw, info = tb.read_bundle(gan.checkpoint_prefix)
  assert int(raw["G_optimizer/iter/.ATTRIBUTES/VARIABLE_VALUE"]) == 1
  assert raw["G/layer_with_weights-0/kernel/.ATTRIBUTES/VARIABLE_VALUE"].shape == (7, 7, 3, 64)   # HWIO
  assert info["object_graph"] is not None
  torch.manual_seed(1)
  gan2 = CycleGAN(parse_args(argv), strategy) if False else None
  args2 = parse_args(argv); args2.global_batch_size = 1
  gan2 = CycleGAN(args2, strategy)
  assert gan2.load_checkpoint()
  assert gan2.extra_state == {"epoch": 3}
  a, b = gan.backend.state_dict(), gan2.backend.state_dict()
  for k in a:
    assert torch.allclose(a[k].double(), b[k].double(), rtol=1e-6, atol=1e-12), k
  r1, r2 = gan.distributed_test_step(x, y), gan2.distributed_test_step(x, y)
  for k in r1:
    assert torch.allclose(r1[k], r2[k], rtol=1e-5, atol=1e-6), k
