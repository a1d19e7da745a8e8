"""Per-kernel summary of an `ncu --csv --page raw` log that covers every launch of one train step:
for each distinct kernel (demangled name, template arguments kept) the number of launches, total /
mean duration and the recipe's roofline metrics (DRAM and tensor-pipe utilisation, bytes moved,
registers, occupancy). Usage: python tools/ncu_kernels.py <raw.csv>"""
import collections
import csv
import re
import sys

rows = list(csv.reader(l for l in open(sys.argv[1], errors="replace") if l.startswith('"')))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}


def pick(*names):
  for n in names:
    for h in hdr:
      if h == n or h.startswith(n):
        return col[h]
  return None


c_name = pick("Kernel Name")
want = collections.OrderedDict([
    ("dur_us", pick("gpu__time_duration.sum")),
    ("dram_pct", pick("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")),
    ("sm_pct", pick("sm__throughput.avg.pct_of_peak_sustained_elapsed")),
    ("tensor_pct", pick("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")),
    ("dram_rd_MB", pick("dram__bytes_read.sum")),
    ("dram_wr_MB", pick("dram__bytes_write.sum")),
    ("l2_MB", pick("lts__t_bytes.sum")),
    ("regs", pick("launch__registers_per_thread")),
    ("occ_pct", pick("sm__warps_active.avg.pct_of_peak_sustained_active")),
    ("grid", pick("launch__grid_size")),
])
units = rows[1] if len(rows) > 1 and not rows[1][0].isdigit() else None


def num(v):
  try:
    return float(v.replace(",", ""))
  except ValueError:
    return float("nan")


agg = collections.OrderedDict()
for r in rows[2 if units else 1:]:
  if len(r) <= c_name:
    continue
  name = re.sub(r"\(.*$", "", r[c_name])
  name = re.sub(r"b2::(\(anonymous namespace\)::)?", "", name)
  a = agg.setdefault(name, {"n": 0, "vals": collections.defaultdict(list)})
  a["n"] += 1
  for k, i in want.items():
    if i is not None and i < len(r):
      a["vals"][k].append(num(r[i]))


def scale(k, v):
  u = units[want[k]] if units and want[k] is not None else ""
  if k == "dur_us":
    return v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
  if k.endswith("_MB"):
    return v / {"byte": 1e6, "Kbyte": 1e3, "Mbyte": 1.0, "Gbyte": 1e-3}.get(u, 1e6)
  return v


tot = sum(sum(scale("dur_us", v) for v in a["vals"]["dur_us"]) for a in agg.values())
print(f"# {len(agg)} distinct kernels, {sum(a['n'] for a in agg.values())} launches, {tot / 1e3:.3f} ms summed duration "
      f"(serialised, cold caches under ncu: compare shares, not absolutes)")
print(f"{'kernel':74s} {'n':>4s} {'sum us':>9s} {'%':>5s} {'mean us':>8s} {'dram%':>6s} {'sm%':>6s} {'tens%':>6s} "
      f"{'rd MB':>7s} {'wr MB':>7s} {'L2 MB':>7s} {'regs':>5s} {'occ%':>5s} {'grid':>6s}")
for name, a in sorted(agg.items(), key=lambda kv: -sum(scale("dur_us", v) for v in kv[1]["vals"]["dur_us"])):
  v = a["vals"]
  mean = lambda k: (sum(scale(k, x) for x in v[k]) / len(v[k])) if v[k] else float("nan")
  s = sum(scale("dur_us", x) for x in v["dur_us"])
  print(f"{name[:74]:74s} {a['n']:4d} {s:9.1f} {100 * s / tot:5.1f} {mean('dur_us'):8.1f} {mean('dram_pct'):6.1f} "
        f"{mean('sm_pct'):6.1f} {mean('tensor_pct'):6.1f} {mean('dram_rd_MB'):7.1f} {mean('dram_wr_MB'):7.1f} "
        f"{mean('l2_MB'):7.1f} {mean('regs'):5.0f} {mean('occ_pct'):5.1f} {mean('grid'):6.0f}")
