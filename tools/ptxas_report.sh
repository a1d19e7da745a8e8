#!/bin/bash
# registers / spills of every kernel of one .cu (cross-compiles, no GPU needed)
#   tools/ptxas_report.sh tf2_cyclegan_b200/csrc/elementwise.cu [filter-regex]
src=$1; filter=${2:-.}
dir=$(dirname "$src")
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr \
  -Xptxas -v -I"$dir" -I"$(python -c 'import sysconfig;print(sysconfig.get_paths()["include"])')" \
  -I"$(python -c 'import pybind11;print(pybind11.get_include())')" -c "$src" -o /tmp/ptxas_report.o 2>&1 |
  awk -F"'" '/Compiling entry function/ {name=$2} /spill stores/ {spill=$0} /Used [0-9]+ registers/ {print name "|" $0 "|" spill}' |
  while IFS='|' read -r n u s; do echo "$(echo "$n" | c++filt | sed -E 's/b2::\(anonymous namespace\):://; s/\(b2::.*//; s/^void //') | $(echo "$u" | sed 's/ptxas info    : //') |$s"; done |
  grep -E "$filter"
