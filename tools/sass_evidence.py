"""SASS evidence that the kernels are Blackwell-native (cuobjdump -sass on the built extension):

  profiles/sass_evidence.txt      per-kernel counts of the tcgen05 / TMA / TMEM / multimem / bulk-copy mnemonics
  profiles/sass/<kernel>.sass     FULL listing of every kernel that a train step (bf16, 1..8 GPUs) launches
                                  (instruction text only; the encoding column is dropped)

    python tools/sass_evidence.py [path/to/_C*.so]
"""
import collections
import os
import re
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "tf2_cyclegan_b200/csrc/_C.cpython-312-x86_64-linux-gnu.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCBAR|UTMALDG|UTMASTG|UTMAREDG|UBLKCP|UBLKRED|LDTM|STTM|UTCATOMSWS|LDGMC|STGMC|REDGMC|"
                 r"SYNCS|HMMA|HGMMA|MULTIMEM|LDGSTS|ATOMG|REDG|RED|UCGABAR_ARV|UCGABAR_WAIT|MEMBAR)[\w\.]*")
# kernels of the default step plan (template variants that the launchers can pick for the model's shapes)
HOT = re.compile(r"conv_gemm_ts_kernel|conv_gemm2_kernel|conv_gemm_kernel|wgrad_gemm2_kernel|wgrad_gemm_kernel|"
                 r"wgrad_multi_kernel|in_apply_stream_kernel|in_bwd_reduce_stream_kernel|in_bwd_apply_stream_kernel|"
                 r"allreduce_adam_kernel|act_bwd_kernel|img_prep_kernel|l1_loss_kernel|mse_loss_kernel|tap_gather_kernel|"
                 r"adam_kernel|d4_dgrad_kernel|d4_wgrad_kernel|mask_grad_kernel|gather_bf16|scatter_add_kernel|step_increment")
cur = None
counts = collections.OrderedDict()
text = collections.OrderedDict()
for line in out.splitlines():
  m = re.search(r"Function : (\S+)", line)
  if m:
    cur = m.group(1)
    counts[cur] = collections.Counter()
    text[cur] = []
    continue
  if cur is None:
    continue
  m2 = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?)\s*/\* 0x[0-9a-f]+ \*/", line)
  if m2:
    text[cur].append(f"/*{m2.group(1)}*/ {m2.group(2)}")
  m = pat.search(line)
  if m:
    counts[cur][m.group(0).split(".")[0] + ("." + ".".join(m.group(0).split(".")[1:3]) if "." in m.group(0) else "")] += 1


def demangle(fn):
  d = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
  d = re.sub(r"b2::(\(anonymous namespace\)::)?", "", d)
  return re.sub(r"\(.*$", "", re.sub(r"^void ", "", d))


lines = [f"# SASS mnemonics per kernel in {os.path.relpath(so, root)}",
         "# tcgen05.mma -> UTCHMMA / UTCQMMA, tcgen05.ld -> LDTM, tcgen05.commit -> UTCBAR, cp.async.bulk.tensor -> UTMALDG,",
         "# cp.async.bulk (1-D) -> UBLKCP, cp.reduce.async.bulk -> UBLKRED / UTMAREDG, multimem.ld_reduce -> LDGMC, multimem.st -> STGMC,",
         "# mbarrier -> SYNCS; HMMA would be legacy mma.sync (there is none)"]
sass_dir = os.path.join(root, "profiles", "sass")
os.makedirs(sass_dir, exist_ok=True)
for f in os.listdir(sass_dir):
  os.remove(os.path.join(sass_dir, f))
written = 0
for fn, c in counts.items():
  dem = demangle(fn)
  if c:
    lines.append(f"\n{dem[:150]}   [{len(text[fn])} instructions]")
    lines.append("   " + ", ".join(f"{k} x{v}" for k, v in sorted(c.items())))
  if HOT.search(dem) and text[fn]:
    name = re.sub(r"[^A-Za-z0-9_]+", "_", dem).strip("_")[:120]
    with open(os.path.join(sass_dir, name + ".sass"), "w") as f:
      f.write(f"// {dem}\n// {len(text[fn])} instructions; cuobjdump -sass of {os.path.basename(so)} (sm_100a)\n")
      f.write("\n".join(text[fn]) + "\n")
    written += 1
open(os.path.join(root, "profiles", "sass_evidence.txt"), "w").write("\n".join(lines) + "\n")
print(f"{len(counts)} kernels, {written} full listings under profiles/sass/")
