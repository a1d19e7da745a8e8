"""Run the UMMA descriptor row-shift probe (csrc/shift_probe.cu) and report which variants match."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tf2_cyclegan_b200.ops import native as nv
C = nv.load()
torch.manual_seed(0)
A = torch.randn(144, 64, device="cuda").to(torch.bfloat16)
B = torch.randn(16, 64, device="cuda").to(torch.bfloat16)
out = torch.zeros(16, 128, 16, device="cuda")
ta = nv.tmap(A.data_ptr(), [64, 144], [64], [64, 136])
tb = nv.tmap(B.data_ptr(), [64, 16], [64], [64, 16])
C.shift_probe(ta, tb, out.data_ptr())
for v in range(16):
  s, mode = v & 7, v >> 3
  ref = A[s:s + 128].float() @ B.float().t()
  err = (out[v] - ref).abs().max().item() / ref.abs().max().item()
  print(f"shift {s} base_offset={'s' if mode else '0'}: rel err {err:.4f} {'OK' if err < 1e-2 else 'MISMATCH'}")
