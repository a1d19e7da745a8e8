"""Command line: the reference's five flags verbatim (main.py:405-411) plus
opt-in extras needed for the BASELINE.json configs."""
from __future__ import annotations

import argparse


def build_parser() -> argparse.ArgumentParser:
  parser = argparse.ArgumentParser(description="B200-native CycleGAN trainer")
  # --- reference flags (main.py:407-411)
  parser.add_argument("--output_dir", default="runs")
  parser.add_argument("--epochs", default=200, type=int)
  parser.add_argument("--batch_size", default=1, type=int,
                      help="per-replica batch size; global = × replicas")
  # the reference omits type=int so any CLI value fails its choices check; fixed here
  parser.add_argument("--verbose", default=1, type=int, choices=[0, 1, 2])
  parser.add_argument("--clear_output_dir", action="store_true")
  # --- extras
  parser.add_argument("--image_size", default=256, type=int)
  parser.add_argument("--num_residual_blocks", default=9, type=int)
  parser.add_argument("--dtype", default="bf16", choices=["fp32", "bf16", "fp8", "mxfp8"],
                      help="compute dtype of the native backend (oracle: fp32/bf16); fp8 / mxfp8: residual-block "
                           "forward convs in E4M3 with per-tensor / per-32-channel (MX) scales")
  parser.add_argument("--backend", default="auto", choices=["auto", "native", "oracle"])
  parser.add_argument("--device", default=None, choices=["cpu", "cuda"])
  parser.add_argument("--data_dir", default=None,
                      help="TFDS data dir holding cycle_gan/horse2zebra/<version>/ (tfrecord shards), or a "
                           "folder with trainA/ trainB/ testA/ testB/ images (default: synthetic)")
  parser.add_argument("--synthetic", action="store_true",
                      help="force the synthetic horse2zebra stand-in")
  parser.add_argument("--synthetic_size", default=None, type=int,
                      help="number of synthetic train images per domain (default 1067)")
  parser.add_argument("--seed", default=1234, type=int)
  parser.add_argument("--resume_epoch", action="store_true",
                      help="continue the epoch counter stored in the checkpoint")
  parser.add_argument("--fused_allreduce", default="auto", choices=["auto", "on", "off"],
                      help="fused peer-memory reduce+Adam kernel vs NCCL baseline path")
  parser.add_argument("--checkpoint_format", default="native", choices=["native", "tensorbundle"],
                      help="container written by save_checkpoint: own (JSON index, atomic) or the "
                           "reference's TensorFlow TensorBundle; loading auto-detects either")
  return parser


def parse_args(argv=None):
  args = build_parser().parse_args(argv)
  if args.synthetic:
    args.data_dir = None
  if args.device == "cpu" or args.backend == "oracle":
    if args.dtype == "fp8":
      raise SystemExit("--dtype fp8 needs the native backend (the oracle emulates --dtype mxfp8 only)")
  return args


def main(argv=None):
  from .runner import main as run
  return run(parse_args(argv))


if __name__ == "__main__":
  main()
