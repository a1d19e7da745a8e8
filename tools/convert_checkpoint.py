#!/usr/bin/env python
"""Convert a checkpoint between this framework's container and the reference's TensorFlow
TensorBundle (both live at ``<prefix>.index`` + ``<prefix>.data-00000-of-00001``).

  python tools/convert_checkpoint.py runs/checkpoints/checkpoint out/checkpoint --to tensorbundle
  python tools/convert_checkpoint.py tf_run/checkpoints/checkpoint runs/checkpoints/checkpoint --to native
  python tools/convert_checkpoint.py runs/checkpoints/checkpoint --list
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from tf2_cyclegan_b200.utils import checkpoint as ckpt  # noqa: E402
from tf2_cyclegan_b200.utils import tensorbundle as tb  # noqa: E402


def main():
  ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
  ap.add_argument("src")
  ap.add_argument("dst", nargs="?")
  ap.add_argument("--to", choices=["native", "tensorbundle"], default="tensorbundle")
  ap.add_argument("--list", action="store_true", help="print keys / dtypes / shapes of src and exit")
  a = ap.parse_args()
  kind = "tensorbundle" if tb.is_bundle(a.src) else "native"
  tensors, extra = ckpt.load(a.src)
  print(f"{a.src}: {kind}, {len(tensors)} tensors, "
        f"{sum(t.numel() * t.element_size() for t in tensors.values()) / 1e6:.1f} MB, extra={extra}")
  if a.list or not a.dst:
    for k in sorted(tensors):
      print(f"  {k}  {str(tensors[k].dtype).replace('torch.', '')}{list(tensors[k].shape)}")
    return
  ckpt.save(a.dst, tensors, extra=extra, fmt=a.to)
  print(f"wrote {a.dst} ({a.to})")


if __name__ == "__main__":
  main()
