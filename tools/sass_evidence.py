"""SASS evidence that the kernels are Blackwell-native: per-kernel counts of the tcgen05 / TMA /
TMEM / multimem machine instructions (cuobjdump -sass on the built extension)."""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "tf2_cyclegan_b200/csrc/_C.cpython-312-x86_64-linux-gnu.so"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCBAR|UTMALDG|UTMASTG|UTMAREDG|UBLKCP|UBLKRED|LDTM|STTM|UTCATOMSWS|LDGMC|STGMC|REDGMC|"
                 r"SYNCS|HMMA|HGMMA|MULTIMEM|LDGSTS|ATOMG|REDG|RED|UCGABAR_ARV|UCGABAR_WAIT|MEMBAR)[\w\.]*")
cur = None
counts = collections.OrderedDict()
for line in out.splitlines():
  m = re.search(r"Function : (\S+)", line)
  if m:
    cur = m.group(1)
    counts[cur] = collections.Counter()
    continue
  if cur is None:
    continue
  m = pat.search(line)
  if m:
    counts[cur][m.group(0).split(".")[0] + ("." + ".".join(m.group(0).split(".")[1:3]) if "." in m.group(0) else "")] += 1
print(f"# SASS mnemonics per kernel in {so}")
print("# tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, tcgen05.commit -> UTCBAR, cp.async.bulk.tensor -> UTMALDG,")
print("# cp.reduce.async.bulk -> UBLKRED/UTMAREDG, multimem.ld_reduce -> LDGMC, mbarrier -> SYNCS; HMMA would be legacy mma.sync")
for fn, c in counts.items():
  if not c:
    continue
  dem = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()[:150]
  print(f"\n{dem}")
  print("   " + ", ".join(f"{k} x{v}" for k, v in sorted(c.items())))
