"""config.ENGINE: the single configuration object (replaces the former B2_* environment switches)."""
import subprocess
import sys

import pytest

from tf2_cyclegan_b200.config import EngineConfig


def test_defaults_and_string_updates():
  c = EngineConfig()
  assert c.use_graph and c.lanes == 4 and c.ar_mode == "pull" and c.kernel_debug == 0 and c.nvls == "auto"
  c.update_from_string("use_graph=0, lanes=2,ar_mode=push,wgrad_ctas=96,nvtx=true")
  assert (c.use_graph, c.lanes, c.ar_mode, c.wgrad_ctas, c.nvtx) == (False, 2, "push", 96, True)
  for bad in ("lanes=3", "ar_mode=ring", "nope=1", "use_graph=maybe", "lanes", "kernel_debug=9"):
    with pytest.raises(ValueError):
      EngineConfig().update_from_string(bad)


def test_environment_is_read_once_through_b2_config():
  code = "from tf2_cyclegan_b200.config import ENGINE; print(ENGINE.use_graph, ENGINE.lanes)"
  out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True,
                       env={"B2_CONFIG": "use_graph=0,lanes=1", "PATH": "/usr/bin:/bin"}, cwd=".")
  assert out.stdout.split() == ["False", "1"], out
  # no other B2_* variable is consulted anywhere in the package
  import os
  import re
  hits = []
  for root, _, files in os.walk("tf2_cyclegan_b200"):
    for f in files:
      if f.endswith((".py", ".cu", ".cuh", ".cpp")):
        for m in re.finditer(r'(?:getenv|environ(?:\.get)?)\s*[\(\[]\s*"(B2_[A-Z0-9_]+)"', open(os.path.join(root, f)).read()):
          hits.append((f, m.group(1)))
  assert {h[1] for h in hits} <= {"B2_CONFIG", "B2_NO_HOST_BUILD"}, hits
