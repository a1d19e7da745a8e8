"""CPU smoke tests of the offline tools."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_critical_path_tool_runs_on_a_profile():
  prof = os.path.join(ROOT, "profiles", "r2", "step_profile.txt")
  r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "critical_path.py"), prof],
                     capture_output=True, text=True, timeout=600, cwd=ROOT)
  assert r.returncode == 0, r.stderr[-800:]
  line = next(l for l in r.stdout.splitlines() if "critical path" in l and "serialised" in l)
  total = float(line.split("serialised op time")[1].split("ms")[0])
  cp = float(line.split("dependency critical path")[1].split("ms")[0])
  assert 0 < cp < total            # the plan has parallel lanes
  assert total > 10                # the profile's op classes were matched to plan ops


def test_convert_checkpoint_tool(tmp_path):
  import torch
  sys.path.insert(0, ROOT)
  from tf2_cyclegan_b200.utils import checkpoint as ckpt
  from tf2_cyclegan_b200.utils import tensorbundle as tb
  src = str(tmp_path / "a" / "checkpoint")
  ckpt.save(src, {"G/layer_with_weights-0/kernel/.ATTRIBUTES/VARIABLE_VALUE": torch.randn(3, 3, 8, 8)},
            extra={"epoch": 2})
  dst = str(tmp_path / "b" / "checkpoint")
  r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "convert_checkpoint.py"), src, dst,
                      "--to", "tensorbundle"], capture_output=True, text=True, timeout=300, cwd=ROOT)
  assert r.returncode == 0, r.stderr[-800:]
  assert tb.is_bundle(dst)
  a, ea = ckpt.load(src)
  b, eb = ckpt.load(dst)
  assert ea == eb == {"epoch": 2} and all(torch.equal(a[k], b[k]) for k in a)


def test_shell_tools_parse():
  """The GPU-side evidence / probe scripts are only ever run on a box: at least keep them syntactically valid."""
  import glob
  scripts = sorted(glob.glob(os.path.join(ROOT, "tools", "*.sh")))
  assert len(scripts) >= 6
  for s in scripts:
    r = subprocess.run(["bash", "-n", s], capture_output=True, text=True)
    assert r.returncode == 0, (s, r.stderr)


def test_quality_run_summarises_a_saved_run():
  saved = os.path.join(ROOT, "profiles", "r2", "quality_run_mxfp8.json")
  r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "quality_run.py"), "--summarize", saved],
                     capture_output=True, text=True, timeout=300, cwd=ROOT)
  assert r.returncode == 0, r.stderr[-800:]
  assert "error/MAE(X, F(G(X)))" in r.stdout and "first half of the run" in r.stdout
