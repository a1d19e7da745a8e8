"""Data pipeline semantics, TensorBoard event files, checkpoint container, CLI."""
import os
import types

import numpy as np
import pytest
import torch

from tf2_cyclegan_b200.cli import build_parser
from tf2_cyclegan_b200.data import pipeline as P
from tf2_cyclegan_b200.parallel.strategy import Strategy
from tf2_cyclegan_b200.utils import checkpoint as ckpt
from tf2_cyclegan_b200.utils.dicts import append_dict
from tf2_cyclegan_b200.utils.summary import Summary, crc32c


def test_cli_flags_match_reference():
  a = build_parser().parse_args([])
  assert (a.output_dir, a.epochs, a.batch_size, a.verbose, a.clear_output_dir) == ("runs", 200, 1, 1, False)
  assert build_parser().parse_args(["--verbose", "0"]).verbose == 0
  with pytest.raises(SystemExit):
    build_parser().parse_args(["--verbose", "3"])


def test_tf_shuffle_is_permutation_with_bounded_displacement():
  order = P.tf_shuffle_order(1000, 256, np.random.default_rng(0))
  assert sorted(order) == list(range(1000))
  # an element can only be emitted once it entered the 256-slot buffer
  assert all(v < i + 256 for i, v in enumerate(order))
  assert order != list(range(1000))


def test_shard_bounds_uneven_and_empty():
  s = Strategy(device="cpu")
  s.world_size, s.rank = 4, 0
  shares = []
  for r in range(4):
    s.rank = r
    shares.append(s.shard_bounds(6))
  assert shares == [(0, 2), (2, 4), (4, 5), (5, 6)]
  s.rank = 3
  assert s.shard_bounds(3) == (3, 3)


def _args(**kw):
  d = dict(global_batch_size=4, image_size=32, seed=1, synthetic_size=10, data_dir=None)
  d.update(kw)
  return types.SimpleNamespace(**d)


def test_datasets_steps_partial_batch_and_frozen_augmentation():
  args = _args()
  st = Strategy(device="cpu")
  train_ds, test_ds, plot_ds = P.get_datasets(args, st)
  assert args.train_steps == 3 and len(train_ds) == 3       # ceil(10 / 4)
  sizes = [x.shape[0] for x, _ in train_ds]
  assert sizes == [4, 4, 2]                                   # no drop_remainder
  e1 = torch.cat([x for x, _ in train_ds]).sum(dim=(1, 2, 3)).sort().values
  e2 = torch.cat([x for x, _ in train_ds]).sum(dim=(1, 2, 3)).sort().values
  assert torch.allclose(e1, e2)                               # cache() after random map
  x, y = next(iter(plot_ds))
  assert x.shape == (1, 32, 32, 3) and len(plot_ds) == 1   # min(5, num_test=1)
  assert x.min() >= -1 and x.max() <= 1


def test_folder_source(tmp_path):
  from PIL import Image
  for split in P.H2Z_SPLITS:
    os.makedirs(tmp_path / split)
    for i in range(3):
      Image.fromarray(np.random.randint(0, 255, (40, 50, 3), dtype=np.uint8)).save(tmp_path / split / f"{i}.png")
  args = _args(data_dir=str(tmp_path), global_batch_size=2)
  train_ds, test_ds, _ = P.get_datasets(args, Strategy(device="cpu"))
  assert args.train_steps == 2 and args.test_steps == 2
  x, y = next(iter(test_ds))
  assert x.shape == (2, 32, 32, 3)


def test_resize_matches_tf_half_pixel():
  img = torch.arange(4.0).reshape(2, 2, 1)
  out = P.resize_bilinear(img, (4, 4))[:, :, 0]
  # TF2 half-pixel bilinear of [[0,1],[2,3]] → first row 0, .25, .75, 1
  assert torch.allclose(out[0], torch.tensor([0.0, 0.25, 0.75, 1.0]))


def test_crc32c_known_vector():
  assert crc32c(b"123456789") == 0xE3069283


def test_summary_event_files_and_tags(tmp_path):
  s = Summary(str(tmp_path))
  s.scalar("loss_G/loss", 1.25, step=2, training=True)
  s.scalar("elapse", 3.0, step=2)
  imgs = [np.random.randint(0, 255, (2, 16, 16, 3), dtype=np.uint8)] * 3
  s.image_cycle("X_cycle", imgs, ["X", "G(X)", "F(G(X))"], step=0)
  s.close()
  ea_mod = pytest.importorskip("tensorboard.backend.event_processing.event_accumulator")
  tr = ea_mod.EventAccumulator(str(tmp_path), size_guidance={"tensors": 0})
  tr.Reload()
  assert tr.Tags()["tensors"] == ["loss_G/loss"]
  ev = tr.Tensors("loss_G/loss")[0]
  assert ev.step == 2 and abs(ev.tensor_proto.float_val[0] - 1.25) < 1e-6
  te = ea_mod.EventAccumulator(str(tmp_path / "test"), size_guidance={"tensors": 0})
  te.Reload()
  assert set(te.Tags()["tensors"]) == {"elapse", "X_cycle/sample_#000", "X_cycle/sample_#001"}
  img = te.Tensors("X_cycle/sample_#000")[0].tensor_proto.string_val
  assert img[0] == b"1080" and img[1] == b"390" and img[2][:4] == b"\x89PNG"


def test_checkpoint_roundtrip_and_torn_write(tmp_path):
  prefix = str(tmp_path / "checkpoints" / "checkpoint")
  assert not ckpt.exists(prefix)
  t = {"G/layer_with_weights-0/kernel/.ATTRIBUTES/VARIABLE_VALUE": torch.randn(7, 7, 3, 64),
       "G_optimizer/iter/.ATTRIBUTES/VARIABLE_VALUE": torch.tensor(5.0, dtype=torch.float64),
       "half": torch.randn(5).to(torch.bfloat16)}
  ckpt.save(prefix, t, extra={"epoch": 3})
  assert ckpt.exists(prefix) and os.path.exists(prefix + ".data-00000-of-00001")
  out, extra = ckpt.load(prefix)
  assert extra["epoch"] == 3
  for k in t:
    assert out[k].dtype == t[k].dtype and torch.equal(out[k], t[k])
  ckpt.save(prefix, t)                          # overwrite in place
  with open(prefix + ".data-00000-of-00001", "ab") as f:
    pass
  # simulate a torn write: data from another save
  with open(prefix + ".data-00000-of-00001", "r+b") as f:
    f.seek(-16, 2)
    f.write(b"0" * 16)
  with pytest.raises(IOError):
    ckpt.load(prefix)


def test_append_dict():
  d = {}
  append_dict(d, {"a": 1})
  append_dict(d, {"a": 2})
  assert d == {"a": [1, 2]}
  append_dict(d, {"a": 3}, replace=True)
  assert d == {"a": 3}


def test_prefetch_iterator_abandoned_early_exits_cleanly(tmp_path):
  """A consumer that stops iterating a prefetching dataset early (exception in the training loop,
  a benchmark that takes K batches) must not leave an un-joined worker thread behind: that
  aborts the interpreter at exit ("terminate called without an active exception")."""
  import subprocess
  import sys
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  code = (
      "import torch\n"
      "from tf2_cyclegan_b200.data.pipeline import PairedBatches\n"
      "from tf2_cyclegan_b200.parallel.strategy import Strategy\n"
      "a = torch.rand(14, 8, 8, 3); b = a.flip(0).contiguous()\n"
      "it = iter(PairedBatches(a, b, 2, Strategy(device='cpu'), distribute=False, prefetch=2))\n"
      "x, y = next(it); x, y = next(it)\n"
      "assert torch.equal(x, a[2:4])\n"
      "print('done')\n")
  r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120,
                     env=dict(os.environ, PYTHONPATH=root))
  assert r.returncode == 0 and "done" in r.stdout, (r.returncode, r.stderr[-500:])


def test_weights_only_checkpoint_resumes_with_fresh_optimizer(tmp_path, capsys):
  """main.py:162-170: ``checkpoint.read()`` never asserts that optimizer slots were found, so a bundle
  without them (weights-only export, or written before the first step) must load: Adam restarts from
  zero with a warning; missing WEIGHTS still raise unless expect_partial."""
  import types
  import pytest
  from tf2_cyclegan_b200.engine.cyclegan import CycleGAN
  from tf2_cyclegan_b200.parallel.strategy import Strategy
  from tf2_cyclegan_b200.utils import checkpoint as ckpt
  args = types.SimpleNamespace(output_dir=str(tmp_path), global_batch_size=1, batch_size=1, image_size=16,
                               num_residual_blocks=1, backend="oracle", dtype="fp32", seed=3)
  gan = CycleGAN(args, Strategy(device="cpu"))
  x = torch.rand(1, 16, 16, 3) * 2 - 1
  gan.train_step(x, x.flip(1))
  sd = gan.backend.state_dict()
  weights = {k: v for k, v in sd.items() if "OPTIMIZER_SLOT" not in k and "_optimizer/" not in k}
  for fmt in ("native", "tensorbundle"):
    ckpt.save(gan.checkpoint_prefix, weights, fmt=fmt)
    fresh = CycleGAN(args, Strategy(device="cpu"))
    assert fresh.load_checkpoint()
    assert "optimizer entries missing" in capsys.readouterr().out
    k = "G/layer_with_weights-0/kernel/.ATTRIBUTES/VARIABLE_VALUE"
    assert torch.equal(fresh.backend.state_dict()[k], sd[k])
    assert fresh.backend.opt["G"].iterations == 0
  del weights["F/layer_with_weights-0/kernel/.ATTRIBUTES/VARIABLE_VALUE"]
  ckpt.save(gan.checkpoint_prefix, weights)
  with pytest.raises(KeyError):
    CycleGAN(args, Strategy(device="cpu")).load_checkpoint()
  assert CycleGAN(args, Strategy(device="cpu")).load_checkpoint(expect_partial=True)


def test_torn_tensorbundle_write_is_detected(tmp_path):
  """New data file next to an old index (crash between the two renames): refused, not loaded."""
  import pytest
  from tf2_cyclegan_b200.utils import tensorbundle as tb
  prefix = str(tmp_path / "checkpoint")
  tb.write_bundle(prefix, {"a": torch.arange(10.0)})
  old_index = open(prefix + ".index", "rb").read()
  tb.write_bundle(prefix, {"a": torch.arange(12.0), "b": torch.ones(3)})
  open(prefix + ".index", "wb").write(old_index)
  with pytest.raises(IOError):
    tb.read_bundle(prefix)
