"""DP logic on the gloo fake backend: N-rank run == 1-rank run with N× batch."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    return s.getsockname()[1]


class A:
  num_residual_blocks = 1


def _worker(rank, world, port, x, y, out):
  os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                    MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  torch.set_num_threads(1)
  from tf2_cyclegan_b200.engine.cyclegan import build_specs
  from tf2_cyclegan_b200.engine.oracle_backend import NETS, OracleBackend
  from tf2_cyclegan_b200.parallel.strategy import Strategy
  st = Strategy(device="cpu")
  be = OracleBackend(build_specs(A()), x.shape[0], st, seed=3)
  res = be.train_step(st.shard(x), st.shard(y))
  res = st.reduce_dict(res)
  if rank == 0:
    flat = {n: torch.cat([p.detach().reshape(-1) for p in be.params[n].values()]) for n in NETS}
    out.put(({k: float(v) for k, v in res.items()}, {k: v.numpy() for k, v in flat.items()}))
  st.barrier()
  st.shutdown()


# world 1 / 2 / 4, even shares, uneven shares (3 over 2, 6 over 4) and EMPTY shares
# (2 and 3 samples over 4 ranks: ranks 2-3 / rank 3 get nothing but still join every collective)
@pytest.mark.parametrize("world,batch", [(2, 2), (2, 3), (4, 4), (4, 6), (4, 2), (4, 3)])
def test_dp_equivalence(world, batch):
  torch.manual_seed(0)
  x, y = torch.randn(batch, 16, 16, 3), torch.randn(batch, 16, 16, 3)
  ctx = mp.get_context("spawn")
  results = []
  for w in (1, world):
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, w, port, x, y, q)) for r in range(w)]
    for p in procs:
      p.start()
    results.append(q.get())
    for p in procs:
      p.join(120)
      assert p.exitcode == 0
  (r1, w1), (rn, wn) = results
  for k in r1:
    assert abs(r1[k] - rn[k]) < 1e-5 * max(1, abs(r1[k])), k
  for n in w1:
    assert abs(w1[n] - wn[n]).max() < 1e-5, n
