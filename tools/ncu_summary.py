"""Summarise an Nsight Compute report (read with `ncu -i ... --page raw --csv`) into the
metrics the B200 profiling recipe asks for."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_tensor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active", "gpu__dram_throughput",
        "lts__t_bytes.sum", "lts__throughput", "lts__t_sector_hit_rate", "l1tex__throughput",
        "sm__throughput.avg.pct", "sm__warps_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__cluster", "launch__shared_mem_per_block_dynamic", "smsp__cycles_active.avg",
        "sm__cycles_elapsed.max", "sm__cycles_active.avg", "lts__t_sectors_op_read.sum",
        "lts__t_sectors_op_write.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
        "sm__pipe_tensor_op", "smsp__inst_executed.sum", "tensor"]


def main(path):
  out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
  rows = list(csv.reader(out.splitlines()))
  hdr, units = rows[0], rows[1]
  for vals in rows[2:]:
    name = vals[hdr.index("Kernel Name")]
    print(f"== {name[:100]}  grid {vals[hdr.index('Grid Size')]} block {vals[hdr.index('Block Size')]}")
    for h, u, v in zip(hdr, units, vals):
      if any(k in h for k in KEYS):
        print(f"  {h:78s} {v:>18s} {u}")


if __name__ == "__main__":
  for p in sys.argv[1:]:
    main(p)
