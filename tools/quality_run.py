"""Quality evidence on synthetic data: the SAME training run (runner.main: shuffled data, partial last
batch, test pass and MAE metrics every epoch) with the native backend (bf16 kernels) and with the fp32
autograd oracle, from the same seed. Prints / writes the per-epoch curves of the four test MAEs the
reference reports (README.md:37) and of the train totals, and their largest relative deviation.

    python tools/quality_run.py [--epochs 12] [--size 128] [--images 64] [--batch 4] [--out gpurun_out/quality.json]
"""
import argparse
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def report(curves, config, out=None):
  """Per-tag first / last values and the largest relative deviation native vs oracle, over the first
  half of the epochs (where the two trajectories are still the same optimisation path) and over all
  of them (GAN dynamics amplify rounding differences: late epochs are two different, equally valid runs)."""
  rows, worst_early, worst = [], 0.0, 0.0
  for split in ("test", "train"):
    for t in curves["native"][split]:
      n, o = curves["native"][split][t], curves["oracle"][split][t]
      devs = [abs(x - y) / max(abs(y), 0.05) for x, y in zip(n, o)]
      half = max(1, len(devs) // 2)
      worst_early, worst = max(worst_early, max(devs[:half])), max(worst, max(devs))
      rows.append((split, t, n[0], n[-1], o[0], o[-1], max(devs[:half]), max(devs)))
  print(f"{config['epochs']} epochs, {config['images']} synthetic images per domain at {config['size']}x{config['size']}, "
        f"batch {config['batch']}, {config['nres']} res blocks")
  print(f"{'split':6s} {'tag':26s} {'native e0':>10s} {'native last':>12s} {'oracle e0':>10s} {'oracle last':>12s} "
        f"{'dev 1st half':>13s} {'dev all':>9s}")
  for r in rows:
    print(f"{r[0]:6s} {r[1]:26s} {r[2]:10.4f} {r[3]:12.4f} {r[4]:10.4f} {r[5]:12.4f} {100 * r[6]:12.2f}% {100 * r[7]:8.2f}%")
  print(f"largest relative deviation of any per-epoch value: first half of the run {100 * worst_early:.2f}%, "
        f"whole run {100 * worst:.2f}%")
  if out:
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    json.dump({"config": config, "curves": curves, "max_rel_dev_first_half": worst_early, "max_rel_dev": worst},
              open(out, "w"), indent=1)


ap = argparse.ArgumentParser()
ap.add_argument("--epochs", type=int, default=12)
ap.add_argument("--size", type=int, default=128)
ap.add_argument("--images", type=int, default=64)
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--nres", type=int, default=9)
ap.add_argument("--out", default="gpurun_out/quality.json")
ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp8", "mxfp8"], help="compute dtype of the native run")
ap.add_argument("--summarize", default="", help="re-print the table of a saved json (no GPU needed)")
a = ap.parse_args()
if a.summarize:
  saved = json.load(open(a.summarize))
  report(saved["curves"], saved["config"])
  sys.exit(0)

from tensorboard.backend.event_processing import event_accumulator as ea_mod
from tensorboard.util import tensor_util

from tf2_cyclegan_b200.cli import parse_args
from tf2_cyclegan_b200.runner import main

TEST_TAGS = ["error/MAE(X, F(G(X)))", "error/MAE(Y, G(F(Y)))", "error/MAE(X, F(X))", "error/MAE(Y, G(Y))"]
TRAIN_TAGS = ["loss_G/total", "loss_F/total", "loss_X/loss", "loss_Y/loss"]
curves = {}
with tempfile.TemporaryDirectory() as tmp:
  for backend in ("native", "oracle"):
    out = os.path.join(tmp, backend)
    argv = ["--output_dir", out, "--epochs", str(a.epochs), "--image_size", str(a.size),
            "--num_residual_blocks", str(a.nres), "--synthetic_size", str(a.images), "--batch_size", str(a.batch),
            "--backend", backend, "--verbose", "0", "--dtype", a.dtype if backend == "native" else "fp32"]
    main(parse_args(argv))

    def read(logdir, tags):
      acc = ea_mod.EventAccumulator(logdir, size_guidance={"tensors": 0})
      acc.Reload()
      return {t: [float(tensor_util.make_ndarray(e.tensor_proto)) for e in acc.Tensors(t)] for t in tags}

    curves[backend] = {"test": read(os.path.join(out, "test"), TEST_TAGS), "train": read(out, TRAIN_TAGS)}

report(curves, vars(a), a.out)
