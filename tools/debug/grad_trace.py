"""Where does the native backward first diverge from the (bf16-storage-emulating) oracle?
Compares, for discriminator X, the gradient w.r.t. every raw conv output (native dY buffers) and
every layer output with autograd's, layer by layer from the loss down."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from tf2_cyclegan_b200.engine.cyclegan import build_specs
from tf2_cyclegan_b200.engine.native_backend import NativeBackend, NETS
from tf2_cyclegan_b200.engine.oracle_backend import OracleBackend
from tf2_cyclegan_b200.models import oracle_net
from tf2_cyclegan_b200.ops import oracle as O
from tf2_cyclegan_b200.parallel.strategy import Strategy

st = Strategy(device="cuda")
specs = build_specs(types.SimpleNamespace(num_residual_blocks=1))
b, size = 2, 128
nb = NativeBackend(specs, b, st, seed=11, image_size=size, local_batch=b, use_graph=False)
ob = OracleBackend(specs, b, st, seed=11, emulate_bf16_storage=True)
for n in ob.params:
  for p in ob.params[n].values():
    p.data.copy_(p.data.to(torch.bfloat16).float())
torch.manual_seed(3)
x = torch.rand(b, size, size, 3, device="cuda") * 2 - 1
y = torch.rand(b, size, size, 3, device="cuda") * 2 - 1

# ---- oracle: capture gradients at the stored points of network X applied to [fake_x | x]
captured = {}
orig_layer = oracle_net._layer
def traced_layer(l, params, t, rnd=None):
  out = orig_layer(l, params, t, rnd)
  return out
# re-implement X's forward with hooks
def fwd_X(inp, tag):
  t = ob.rnd(inp)
  for j, l in enumerate(specs["X"].layers):
    p = f"layer_with_weights-{l.lw_index}"
    kernel = ob.params["X"][f"{p}/kernel"]
    bias = ob.params["X"].get(f"{p}/bias") if l.bias else None
    yv = O.conv2d(t, kernel, l.stride, "same", bias)
    if l.norm:
      yv = ob.rnd(yv)
      yv.register_hook(lambda g, j=j: captured.__setitem__((tag, j, "draw"), g.detach().clone()))
      captured[(tag, j, "raw")] = yv.detach()
      q = f"layer_with_weights-{l.lw_index + 1}"
      yv = O.instance_norm(yv, ob.params["X"][f"{q}/gamma"], ob.params["X"][f"{q}/beta"])
    if l.act == "leaky":
      yv = O.leaky_relu(yv, 0.2)
    yv = ob.rnd(yv)
    yv.register_hook(lambda g, j=j: captured.__setitem__((tag, j, "dout"), g.detach().clone()))
    t = yv
  return t

fake_x = ob.net("F", y).detach()
L = ob.losses
loss = L.discriminator_loss(fwd_X(x, "real"), fwd_X(fake_x, "fake"))
loss.backward()

# ---- native: run the whole step up to X's backward end (no optimizer), eager
sp = nb.plan_for(b)
nb._load_inputs(sp, x, y)
s = torch.cuda.current_stream().cuda_stream
order = sorted(NETS, key=lambda k: sp.marks[f"{k}_bwd_end"])
pos = 0
for n in order:
  sp.plan.run(pos, sp.marks[f"{n}_bwd_end"], s, True)
  pos = sp.marks[f"{n}_opt_end"]
torch.cuda.synchronize()
X = sp.ex["X"]
rel = lambda a, r: ((a.float() - r.float()).norm() / r.float().norm().clamp_min(1e-20)).item()
print("X layers:", [l.name for l in X.layers])
for j, l in enumerate(X.layers):
  # native batch layout: [fake (b) | real (b)]
  for tag, sl in (("fake", slice(0, b)), ("real", slice(b, 2 * b))):
    if l.norm:
      nat = X.dY[j].interior()[sl] if X.dY[j].halo else X.dY[j].t[sl]
      print(f"layer {j} {l.name:8s} {tag}: d(raw conv out)  rel L2 err {rel(nat, captured[(tag, j, 'draw')]):.4f}"
            f"   | raw fwd err {rel(X.raw[j].t[sl], captured[(tag, j, 'raw')]):.5f}")
    if X.dyt[j] is not None:
      pass
