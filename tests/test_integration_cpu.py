"""BASELINE.json config 1: 64×64 synthetic, 1 ResNet block, world_size=1 on CPU."""
import os

import pytest

from tf2_cyclegan_b200.cli import parse_args
from tf2_cyclegan_b200.runner import main
from tf2_cyclegan_b200.utils import checkpoint as ckpt

TAGS10 = {"loss_G/loss", "loss_G/cycle", "loss_G/identity", "loss_G/total", "loss_F/loss",
          "loss_F/cycle", "loss_F/identity", "loss_F/total", "loss_X/loss", "loss_Y/loss"}
ERR4 = {"error/MAE(X, F(G(X)))", "error/MAE(Y, G(F(Y)))", "error/MAE(X, F(X))", "error/MAE(Y, G(Y))"}


def _run(out, epochs, extra=()):
  argv = ["--output_dir", str(out), "--epochs", str(epochs), "--image_size", "64",
          "--num_residual_blocks", "1", "--synthetic_size", "4", "--device", "cpu",
          "--backend", "oracle", "--dtype", "fp32", "--verbose", "0", *extra]
  return main(parse_args(argv))


def test_two_epochs_checkpoint_resume_and_tags(tmp_path, capsys):
  out = tmp_path / "runs"
  res = _run(out, 2)
  assert set(res) == TAGS10 | ERR4
  prefix = str(out / "checkpoints" / "checkpoint")
  assert ckpt.exists(prefix)
  sd, extra = ckpt.load(prefix)
  assert extra["epoch"] == 1
  assert sd["G_optimizer/iter/.ATTRIBUTES/VARIABLE_VALUE"].item() == 2 * 4  # 4 steps/epoch
  assert "G/layer_with_weights-0/kernel/.ATTRIBUTES/VARIABLE_VALUE" in sd
  assert "X/layer_with_weights-7/bias/.OPTIMIZER_SLOT/X_optimizer/v/.ATTRIBUTES/VARIABLE_VALUE" in sd
  assert sd["G/layer_with_weights-0/kernel/.ATTRIBUTES/VARIABLE_VALUE"].shape == (7, 7, 3, 64)

  ea_mod = pytest.importorskip("tensorboard.backend.event_processing.event_accumulator")
  tr = ea_mod.EventAccumulator(str(out), size_guidance={"tensors": 0}); tr.Reload()
  assert set(tr.Tags()["tensors"]) == TAGS10
  assert [e.step for e in tr.Tensors("loss_G/total")] == [0, 1]
  te = ea_mod.EventAccumulator(str(out / "test"), size_guidance={"tensors": 0}); te.Reload()
  tags = set(te.Tags()["tensors"])
  assert TAGS10 | ERR4 | {"elapse"} <= tags
  assert {f"X_cycle/sample_#{i:03d}" for i in range(1)} <= tags and "Y_cycle/sample_#000" in tags

  # resume: picks the checkpoint up, epoch counter restarts at 0 like the reference
  capsys.readouterr()
  _run(out, 1)
  log = capsys.readouterr().out
  assert "loaded checkpoint from" in log and "Epoch 001/001" in log
  sd2, _ = ckpt.load(prefix)
  assert sd2["G_optimizer/iter/.ATTRIBUTES/VARIABLE_VALUE"].item() == 12


def test_clear_output_dir(tmp_path):
  out = tmp_path / "runs"
  os.makedirs(out)
  (out / "stale.txt").write_text("x")
  _run(out, 1, extra=["--clear_output_dir"])
  assert not (out / "stale.txt").exists()
