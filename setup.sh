#!/bin/sh
# Build the native sm_100a extension in-tree (the reference's setup.sh installs its
# conda/pip dependencies, setup.sh:1-34; here the only build step is nvcc).
set -e
cd "$(dirname "$0")"
PY=${PYTHON:-python}
if ! command -v nvcc >/dev/null 2>&1 && [ ! -x /usr/local/cuda/bin/nvcc ]; then
  echo "nvcc not found: only the CPU (oracle) backend will be available" >&2
  exit 0
fi
$PY tf2_cyclegan_b200/csrc/build.py "$@"
$PY -c "import tf2_cyclegan_b200; print('b200-cyclegan', tf2_cyclegan_b200.__version__, 'ready')"
