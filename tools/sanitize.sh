#!/bin/bash
# compute-sanitizer passes over the native kernels. Run on a GPU box:
#   bash tools/sanitize.sh [memcheck|racecheck|synccheck|initcheck] ...   (default: memcheck synccheck racecheck)
# Logs go to gpurun_out/sanitizer_<tool>.log; the last line of each is the ERROR SUMMARY.
# memcheck runs the whole smoke step (every kernel of the plan, eager launches); the slower tools
# run the single-kernel tests at their smallest shapes.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TOOLS=${*:-memcheck synccheck racecheck}
export B2_CONFIG=use_graph=0
for tool in $TOOLS; do
  log=gpurun_out/sanitizer_${tool}.log
  if [ "$tool" = memcheck ]; then
    target=(python -c "import __graft_entry__ as g; g.smoke()")
  elif [ "$tool" = racecheck ]; then
    # without the MXFP8 kernels: racecheck flags the weight-scale bulk copy (cp.async.bulk, completion through the
    # stage's mbarrier together with the TMA tensor loads it does not track) against the scale warps' reads, and the
    # flood of reports slows the kernel into its own watchdog. synccheck / memcheck do cover them.
    target=(python -m pytest tests/test_native_kernels.py -x -q -m gpu -k "conv3x3 or in_apply or losses")
  else
    target=(python -m pytest tests/test_native_kernels.py -x -q -m gpu -k "conv3x3 or in_apply or losses or mxfp8")
  fi
  timeout ${SANITIZE_TIMEOUT:-900} compute-sanitizer --tool "$tool" --print-limit 20 \
      --log-file "$log" "${target[@]}" > gpurun_out/sanitizer_${tool}.out 2>&1
  echo "== $tool rc=$? $(grep -a 'ERROR SUMMARY' "$log" | tail -1)"
done
