"""The stand-in baseline (baseline/standin.py: stock PyTorch library kernels, NOT the reference) builds
the same architecture and runs the reference's train-step schedule — checked on CPU at a tiny size."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline"))


def test_standin_architecture_and_step():
  import standin
  gen = torch.Generator().manual_seed(0)
  g = standin.generator_params(gen)
  d = standin.discriminator_params(gen)
  assert sum(t.numel() for _, t in g) == 11_383_427 and len(g) == 71
  assert sum(t.numel() for _, t in d) == 2_765_633 and len(d) == 13
  m = standin.StandinCycleGAN(torch.device("cpu"), local_batch=1, global_batch=2, image_size=32, n_res=1)
  # TF-SAME geometry: 32 -> 16 -> 8 -> 4 -> 4 -> 4 patch map; generator is shape preserving
  x = torch.rand(1, 3, 32, 32).to(torch.bfloat16)
  assert standin.discriminator_forward(m.nets["X"].P, x).shape == (1, 1, 4, 4)
  assert standin.generator_forward(m.nets["G"].P, x, 1).shape == (1, 3, 32, 32)
  w0 = m.nets["G"].master.clone()
  xs, ys = torch.rand(1, 32, 32, 3) * 2 - 1, torch.rand(1, 32, 32, 3) * 2 - 1
  l1 = m.train_step(xs, ys).clone()
  l2 = m.train_step(xs, ys).clone()
  assert torch.isfinite(l1).all() and torch.isfinite(l2).all()
  # losses are divided by the GLOBAL batch (2) although only one local sample was fed: identity loss
  # 5 * MAE / 2 with |x - G(x)| <= 2  =>  <= 5
  assert 0 < l1[2].item() <= 5.0
  assert math.isclose(l1[3].item(), (l1[0] + l1[1] + l1[2]).item(), rel_tol=1e-5)
  step = (m.nets["G"].master - w0).abs().max().item()
  assert 0 < step < 2 * 2e-4 * 2.5        # two Keras-Adam steps at lr 2e-4 (first steps are ~lr each)
  assert int(m.nets["G"].step) == 2
  assert torch.equal(m.nets["G"].leaves[0].float(), m.nets["G"].master_views[0].to(torch.bfloat16).float())
