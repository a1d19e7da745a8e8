"""Fused peer-memory all-reduce + Adam kernel vs NCCL all_reduce + Keras-Adam oracle
(SURVEY §4 'distributed (GPU)'). Needs >= 2 GPUs; skipped otherwise."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]

WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["B2_ROOT"])
from tf2_cyclegan_b200.parallel.strategy import Strategy
from tf2_cyclegan_b200.parallel.fused_allreduce import SymmetricBuffers
from tf2_cyclegan_b200.ops import native as nv
from tf2_cyclegan_b200.ops.keras_adam import KerasAdam
st = Strategy(device="cuda")
C = nv.load()
sb = SymmetricBuffers(st)
ok = True
for n in (1024, 8 * 1000 + 8 * 3, 11_400_000 // 8 * 8):
  g, wb = sb.alloc(f"t{n}", n)
  torch.manual_seed(1)                       # identical initial weights on all ranks
  w0 = torch.randn(n, device="cuda")
  class S: pass
  s = S(); s.n = n
  s.w = w0.clone(); s.m = torch.zeros(n, device="cuda"); s.v = torch.zeros(n, device="cuda")
  s.step = torch.zeros(1, dtype=torch.int32, device="cuda")
  plan = C.Plan()
  sb.emit_allreduce_adam(plan, f"t{n}", s, 2e-4, 0.5, 0.9, 1e-7)
  plan.add_step_increment("inc", s.step.data_ptr())
  ref_w = {"w": w0.clone()}
  opt = KerasAdam()
  for it in range(3):
    torch.manual_seed(100 + 10 * it + st.rank)   # different local gradients per rank
    gl = torch.randn(n, device="cuda")
    g.copy_(gl)
    tot = gl.clone(); dist.all_reduce(tot)
    opt.apply_gradients(ref_w, {"w": tot})
    torch.cuda.synchronize(); dist.barrier()
    plan.run(0, -1, torch.cuda.current_stream().cuda_stream, False)
    torch.cuda.synchronize(); dist.barrier()
    full = sb.gather_sharded(s.w)
    e1 = (full - ref_w["w"]).abs().max().item()
    e2 = (wb.float() - ref_w["w"]).abs().max().item()
    gz = g.abs().max().item()
    if not (e1 < 5e-5 and e2 < 2e-2 and gz == 0.0):   # fp32 sum order differs from NCCL
      ok = False
      print(f"rank {st.rank} n={n} it={it}: master err {e1}, bf16 err {e2}, grad residue {gz}")
print(f"rank {st.rank} multicast={sb.multicast_ptrs('t1024')[0] != 0} ok={ok}")
st.barrier(); st.shutdown()
sys.exit(0 if ok else 1)
'''


def _free_port():
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    return s.getsockname()[1]


@pytest.mark.parametrize("nvls", ["auto", "0"])
def test_fused_allreduce_adam_matches_nccl(tmp_path, nvls):
  ngpu = torch.cuda.device_count()
  if ngpu < 2:
    pytest.skip("needs >= 2 GPUs")
  world = 2 if ngpu < 4 else (4 if ngpu < 8 else 8)
  script = tmp_path / "worker.py"
  script.write_text(WORKER)
  env = dict(os.environ, B2_ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))), B2_NVLS=nvls)
  cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
         "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)]
  r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
  print(r.stdout[-3000:], r.stderr[-3000:])
  assert r.returncode == 0
