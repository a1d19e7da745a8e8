"""Pure-PyTorch oracle ops with the reference's exact TensorFlow semantics.

Every op here is the numerical ground truth for the sm_100a kernels in
``csrc/`` and doubles as the CPU backend. Tensors are NHWC (channels-last),
like the reference (no ``data_format`` is passed anywhere in
cyclegan/model.py), and conv kernels use TF layouts:

* Conv2D kernel            ``[kh, kw, cin, cout]``   (HWIO)
* Conv2DTranspose kernel   ``[kh, kw, cout, cin]``

Semantics that differ from PyTorch defaults:

* TF "SAME" padding is asymmetric (extra pixel goes *after*).
* Conv2DTranspose(k3, s2, "same") is the exact adjoint of the SAME conv: the
  full (2n+1) output cropped at the end.
* InstanceNormalization is tfa's: biased variance, eps = 1e-3.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn.functional as F

IN_EPS = 1e-3  # tfa.layers.InstanceNormalization default (cyclegan/model.py:58)


# ----------------------------------------------------------------- padding
def same_pad_amounts(size: int, k: int, s: int) -> Tuple[int, int]:
  """TF SAME padding (before, after) for one spatial dim."""
  out = math.ceil(size / s)
  total = max((out - 1) * s + k - size, 0)
  return total // 2, total - total // 2


def reflect_pad(x: torch.Tensor, p: int) -> torch.Tensor:
  """tf.pad(mode="REFLECT") on H and W of an NHWC tensor (model.py:14-33)."""
  if p == 0:
    return x
  y = F.pad(x.permute(0, 3, 1, 2), (p, p, p, p), mode="reflect")
  return y.permute(0, 2, 3, 1)


# ------------------------------------------------------------- convolutions
def conv2d(x: torch.Tensor,
           kernel: torch.Tensor,
           stride: int = 1,
           padding: str = "valid",
           bias: Optional[torch.Tensor] = None) -> torch.Tensor:
  """tf.keras.layers.Conv2D on NHWC input with an HWIO kernel."""
  kh, kw = kernel.shape[0], kernel.shape[1]
  xc = x.permute(0, 3, 1, 2)
  if padding == "same":
    pt, pb = same_pad_amounts(x.shape[1], kh, stride)
    pl, pr = same_pad_amounts(x.shape[2], kw, stride)
    xc = F.pad(xc, (pl, pr, pt, pb))
  elif padding != "valid":
    raise ValueError(padding)
  w = kernel.permute(3, 2, 0, 1)  # OIHW
  y = F.conv2d(xc, w, bias=bias, stride=stride)
  return y.permute(0, 2, 3, 1)


def conv2d_transpose(x: torch.Tensor,
                     kernel: torch.Tensor,
                     stride: int = 2,
                     bias: Optional[torch.Tensor] = None) -> torch.Tensor:
  """tf.keras.layers.Conv2DTranspose(padding="same") (model.py:114-121).

  Exact adjoint of ``conv2d(..., padding="same")``: compute the full
  transposed convolution, then crop ``same_pad_amounts`` of the *output* size.
  """
  kh, kw = kernel.shape[0], kernel.shape[1]
  n, h, w_, _ = x.shape
  out_h, out_w = h * stride, w_ * stride
  pt, _ = same_pad_amounts(out_h, kh, stride)
  pl, _ = same_pad_amounts(out_w, kw, stride)
  wt = kernel.permute(3, 2, 0, 1)  # [cin, cout, kh, kw]
  y = F.conv_transpose2d(x.permute(0, 3, 1, 2), wt, bias=bias, stride=stride)
  y = y[:, :, pt:pt + out_h, pl:pl + out_w]
  return y.permute(0, 2, 3, 1)


# ------------------------------------------------------------ normalisation
def instance_norm(x: torch.Tensor,
                  gamma: torch.Tensor,
                  beta: torch.Tensor,
                  eps: float = IN_EPS) -> torch.Tensor:
  """tfa InstanceNormalization on NHWC: per-(n, c) stats over H×W."""
  mean = x.mean(dim=(1, 2), keepdim=True)
  var = x.var(dim=(1, 2), keepdim=True, unbiased=False)
  return (x - mean) * torch.rsqrt(var + eps) * gamma + beta


def leaky_relu(x: torch.Tensor, alpha: float = 0.2) -> torch.Tensor:
  return F.leaky_relu(x, negative_slope=alpha)


# ------------------------------------------------------------------- losses
def mae(y_true: torch.Tensor, y_pred: torch.Tensor) -> torch.Tensor:
  """Per-sample mean absolute error over all non-batch axes (main.py:86-89)."""
  return (y_true - y_pred).abs().flatten(1).mean(dim=1)


def mse(y_true: torch.Tensor, y_pred: torch.Tensor) -> torch.Tensor:
  """Per-sample mean squared error (main.py:92-95)."""
  return (y_true - y_pred).square().flatten(1).mean(dim=1)


def bce(y_true: torch.Tensor, y_pred: torch.Tensor,
        from_logits: bool = False) -> torch.Tensor:
  """Per-sample binary cross-entropy (main.py:98-103; defined, never called).

  Keras clips probabilities to [1e-7, 1 - 1e-7] when ``from_logits=False``.
  """
  if from_logits:
    per = F.binary_cross_entropy_with_logits(y_pred, y_true, reduction="none")
  else:
    p = y_pred.clamp(1e-7, 1.0 - 1e-7)
    per = -(y_true * torch.log(p) + (1.0 - y_true) * torch.log(1.0 - p))
  return per.flatten(1).mean(dim=1)
