"""Data-parallel native training (fused peer-memory reduce + Adam inside the CUDA graph) vs the
single-process autograd oracle on the full global batch. Needs >= 2 GPUs."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]

WORKER = r'''
import os, sys, types, torch
sys.path.insert(0, os.environ["B2_ROOT"])
from tf2_cyclegan_b200.engine.cyclegan import build_specs
from tf2_cyclegan_b200.engine.native_backend import NativeBackend
from tf2_cyclegan_b200.engine.oracle_backend import OracleBackend
from tf2_cyclegan_b200.parallel.strategy import Strategy
st = Strategy(device="cuda")
specs = build_specs(types.SimpleNamespace(num_residual_blocks=1))
gbs = 2 * st.world_size
nb = NativeBackend(specs, gbs, st, seed=21, image_size=128, local_batch=2,
                   fused_allreduce=os.environ.get("B2_FUSED", "auto"))
torch.manual_seed(5)
x = torch.rand(gbs, 128, 128, 3, device="cuda") * 2 - 1
y = torch.rand(gbs, 128, 128, 3, device="cuda") * 2 - 1
for _ in range(3):
  res = nb.reduce_results(nb.train_step(st.shard(x), st.shard(y)))
# a partial global batch: one sample -> rank 0 gets it, every other rank has an EMPTY share but still
# joins every cross-GPU launch in the same order (and sees the same reduced losses)
res_p = nb.reduce_results(nb.train_step(st.shard(x[:1]), st.shard(y[:1])))
sd = nb.state_dict()          # collective (gathers the ZeRO-1 shards)
ok = True
def same_on_all_ranks(t):
  lo, hi = t.float().clone(), t.float().clone()
  torch.distributed.all_reduce(lo, op=torch.distributed.ReduceOp.MIN)
  torch.distributed.all_reduce(hi, op=torch.distributed.ReduceOp.MAX)
  return torch.equal(lo, hi)
for k, v in res_p.items():
  if not same_on_all_ranks(v.reshape(1)):
    ok = False; print("partial-batch results differ across ranks:", k)
# fp32 masters of the vector parameters (InstanceNorm gamma / beta, biases) are read by the forward and
# backward kernels on EVERY rank: they must be current and bit-identical everywhere (fused path: the
# one-shot tail of the flat buffer; it used to be stale on non-owner ranks)
for n, stn in nb.state.items():
  lo = stn.layout.vec_lo
  for name, t in (("w", stn.w[lo:]), ("m", stn.m[lo:]), ("v", stn.v[lo:])):
    if not same_on_all_ranks(t):
      ok = False; print("vector parameters diverge across ranks:", n, name)
  if not torch.equal(stn.wb[lo:], stn.w[lo:].to(torch.bfloat16)):
    ok = False; print("bf16 copy of the vector tail is stale:", n)
# replica consistency: the bf16 compute copy and every derived (re-laid-out) weight tensor must be
# identical on all ranks and current (the fp32 master is sharded, so they may only come from wb)
for n, stn in nb.state.items():
  tensors = [("wb", stn.wb)] + [(k, d) for k, (d, _) in stn.derived.items()]
  for name, t in tensors:
    lo, hi = t.float().clone(), t.float().clone()
    torch.distributed.all_reduce(lo, op=torch.distributed.ReduceOp.MIN)
    torch.distributed.all_reduce(hi, op=torch.distributed.ReduceOp.MAX)
    if not torch.equal(lo, hi):
      ok = False; print("replicas diverge:", n, name, (hi - lo).abs().max().item())
  for k, (d, idx) in stn.derived.items():
    src = torch.cat([stn.wb, stn.wb.new_zeros(1)])
    want = src[idx.long().clamp_min(-1)].view_as(d)
    if not torch.equal(want, d):
      ok = False; print("stale derived weights:", n, k, (want.float() - d.float()).abs().max().item())
if st.rank == 0:
  one = types.SimpleNamespace(world_size=1, rank=0, device=st.device,
                              all_reduce_sum=lambda t: t)
  ob = OracleBackend(specs, gbs, one, seed=21)
  for n in ob.params:
    for p in ob.params[n].values():
      p.data.copy_(p.data.to(torch.bfloat16).float())
  for _ in range(3):
    ro = ob.train_step(x, y)
  for k in ro:
    a, b = float(res[k]), float(ro[k])
    if abs(a - b) > 5e-2 * max(abs(b), 0.05):
      ok = False; print("loss mismatch", k, a, b)
  ro = ob.train_step(x[:1], y[:1])
  for k in ro:
    a, b = float(res_p[k]), float(ro[k])
    if abs(a - b) > 5e-2 * max(abs(b), 0.05):
      ok = False; print("partial-batch loss mismatch", k, a, b)
  osd = ob.state_dict()
  for k in ("G/layer_with_weights-2/kernel/.ATTRIBUTES/VARIABLE_VALUE",
            "X/layer_with_weights-1/kernel/.ATTRIBUTES/VARIABLE_VALUE",
            "F/layer_with_weights-3/gamma/.ATTRIBUTES/VARIABLE_VALUE"):
    d = (sd[k] - osd[k]).abs().max().item()
    if d > 2e-3:
      ok = False; print("weight mismatch", k, d)
  print("collective =", nb.collective(), "ok =", ok)
flag = torch.tensor([1 if ok else 0], device="cuda")
torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
st.barrier(); st.shutdown()
sys.exit(0 if flag.item() == 1 else 1)
'''


def _free_port():
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    return s.getsockname()[1]


@pytest.mark.parametrize("fused", ["auto", "off"])
def test_native_dp_matches_oracle(tmp_path, fused):
  if torch.cuda.device_count() < 2:
    pytest.skip("needs >= 2 GPUs")
  script = tmp_path / "worker.py"
  script.write_text(WORKER)
  env = dict(os.environ, B2_ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))), B2_FUSED=fused)
  cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)]
  r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
  print(r.stdout[-3000:], r.stderr[-3000:])
  assert r.returncode == 0


def test_cli_two_ranks_native(tmp_path):
  """torchrun main.py on 2 GPUs: runner + DP sharding + rank-0-only IO + sharded-state checkpoint."""
  if torch.cuda.device_count() < 2:
    pytest.skip("needs >= 2 GPUs")
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  out = tmp_path / "runs"
  cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(root, "main.py"),
         "--output_dir", str(out), "--epochs", "2", "--image_size", "128", "--num_residual_blocks", "1",
         "--synthetic_size", "7", "--batch_size", "2", "--backend", "native", "--verbose", "0"]
  r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=root)
  print(r.stdout[-2500:], r.stderr[-2500:])
  assert r.returncode == 0
  assert "Number of devices: 2" in r.stdout and "saved checkpoint" in r.stdout
  sys.path.insert(0, root)
  from tf2_cyclegan_b200.utils import checkpoint as ckpt
  sd, extra = ckpt.load(str(out / "checkpoints" / "checkpoint"))
  assert extra["epoch"] == 1
  # 7 train images, global batch 4 -> 2 steps / epoch (the second one partial: 3 samples over 2 ranks)
  assert sd["G_optimizer/iter/.ATTRIBUTES/VARIABLE_VALUE"].item() == 4
  m = sd["G/layer_with_weights-2/kernel/.OPTIMIZER_SLOT/G_optimizer/m/.ATTRIBUTES/VARIABLE_VALUE"]
  assert m.abs().max().item() > 0
  assert len([f for f in os.listdir(out) if f.startswith("events.out.tfevents")]) == 1   # chief only
