"""Build the native extension in-tree with nvcc for sm_100a (no JIT cache: the
built `_C*.so` travels to the GPU box with the repo snapshot)."""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["conv_gemm.cu", "conv_gemm2.cu", "conv_gemm_ts.cu", "conv_gemm_mx.cu", "wgrad_gemm.cu", "wgrad_multi.cu", "elementwise.cu", "allreduce_adam.cu", "plan.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def so_path() -> str:
  return os.path.join(HERE, "_C" + sysconfig.get_config_var("EXT_SUFFIX"))


def host_so_path() -> str:
  return os.path.join(HERE, "_host.so")


def build_host(force: bool = False) -> str:
  """Host-only helpers (crc32c, image preprocessing, batch gather): plain C ABI, g++."""
  src, target = os.path.join(HERE, "hostlib.cpp"), host_so_path()
  if force or _stale(target, [src]):
    cmd = [os.environ.get("CXX", "g++"), "-O3", "-std=c++17", "-msse4.2", "-fPIC", "-shared",
           "-pthread", src, "-o", target]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
      raise RuntimeError("g++ failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
  return target


def _stale(target: str, deps) -> bool:
  if not os.path.exists(target):
    return True
  t = os.path.getmtime(target)
  return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
  import pybind11
  build_host(force)
  nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
  headers = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cuh", ".h"))]
  objdir = os.path.join(HERE, "build")
  os.makedirs(objdir, exist_ok=True)
  inc = ["-I" + sysconfig.get_paths()["include"], "-I" + pybind11.get_include(), "-I" + HERE]
  flags = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
           "-Xptxas", "-v" if verbose else "-O3"]
  objs, jobs = [], []
  for src in SOURCES:
    s = os.path.join(HERE, src)
    o = os.path.join(objdir, src.replace(".cu", ".o"))
    objs.append(o)
    if force or _stale(o, [s] + headers):
      jobs.append([nvcc, *ARCH, *flags, *inc, "-c", s, "-o", o])

  def run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
      raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return r.stdout + r.stderr

  with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
    logs = list(ex.map(run, jobs))
  if verbose:
    print("\n".join(logs))
  target = so_path()
  if jobs or force or _stale(target, objs):
    run([nvcc, *ARCH, "-shared", "-cudart", "shared", *objs, "-o", target])
  return target


if __name__ == "__main__":
  print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
