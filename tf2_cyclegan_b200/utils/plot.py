"""Cycle figures (cyclegan/utils.py:61-98,112-144) rendered with Pillow.

matplotlib is not installed in the target image; the reference's figure is a
1×3 row of images with titles, figsize (9, 3.25) in at dpi 120, no ticks,
wspace 0.02 — reproduced here as a 1080×390 RGBA canvas.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np

from .dicts import append_dict


def cycle_figure(panels: Sequence[np.ndarray], labels: Sequence[str], dpi: int = 120):
  from PIL import Image, ImageDraw, ImageFont
  W, H = int(9 * dpi), int(3.25 * dpi)
  canvas = Image.new("RGBA", (W, H), (255, 255, 255, 255))
  draw = ImageDraw.Draw(canvas)
  try:
    font = ImageFont.load_default(size=max(12, dpi // 7))
  except TypeError:  # older Pillow
    font = ImageFont.load_default()
  title_h = int(0.30 * dpi)
  gap = int(0.02 * W / 3)
  side = min((W - 4 * gap) // 3, H - title_h - gap)
  x0 = (W - 3 * side - 2 * gap) // 2
  for i, (img, label) in enumerate(zip(panels, labels)):
    arr = np.asarray(img)
    if arr.ndim == 3 and arr.shape[-1] == 1:
      arr = arr[..., 0]
    tile = Image.fromarray(arr.astype(np.uint8)).convert("RGBA").resize((side, side), Image.NEAREST)
    left = x0 + i * (side + gap)
    canvas.paste(tile, (left, title_h))
    tw = draw.textlength(label, font=font)
    draw.text((left + (side - tw) / 2, title_h * 0.15), label, fill=(0, 0, 0, 255), font=font)
  return canvas


def to_uint8(images: np.ndarray) -> np.ndarray:
  """Scale [-1, 1] images back to [0, 255] (utils.py:129-130)."""
  return ((images + 1) * 127.5).astype(np.uint8)


def plot_cycle(plot_ds, gan, summary, epoch: int):
  """Run ``gan.cycle_step`` on the plot pairs and write X_cycle / Y_cycle panels
  to the test writer (utils.py:112-144)."""
  samples = {}
  for x, y in plot_ds:
    fake_x, fake_y, cycle_x, cycle_y = gan.cycle_step(x, y, training=False)
    append_dict(samples, {"x": x, "y": y, "fake_x": fake_x, "fake_y": fake_y,
                          "cycle_x": cycle_x, "cycle_y": cycle_y})
  for key, images in samples.items():
    arr = np.concatenate([np.asarray(i.detach().float().cpu()) if hasattr(i, "detach")
                          else np.asarray(i) for i in images], axis=0)
    samples[key] = to_uint8(arr)
  summary.image_cycle(tag="X_cycle",
                      images=[samples["x"], samples["fake_y"], samples["cycle_x"]],
                      labels=["X", "G(X)", "F(G(X))"], step=epoch, training=False)
  summary.image_cycle(tag="Y_cycle",
                      images=[samples["y"], samples["fake_x"], samples["cycle_y"]],
                      labels=["Y", "F(Y)", "G(F(Y))"], step=epoch, training=False)
