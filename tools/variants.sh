#!/bin/bash
# usage: tools/variants.sh "<-D flags variant 1>" "<variant 2>" ...   (runs on the GPU box; rebuilds and times each)
# each variant is timed with ring 1 (per-kernel ms) and, unless QUICK=1, with ring 3 (pipelined pairs/s)
cd "$(dirname "$0")/.."
for v in "$@"; do
  make -s -C turbo_metrics_b200/csrc -B NVCCFLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC --expt-relaxed-constexpr $v" >/dev/null 2>&1 || { echo "BUILD FAILED: $v"; continue; }
  echo "== $v"
  timeout 120 python tools/quick_time.py 3840 2160 16 8 1 96 16 | grep "rep 1"
  [ -n "$QUICK" ] || timeout 120 python tools/quick_time.py 3840 2160 16 8 3 192 32 | grep "rep 1"
done
make -s -C turbo_metrics_b200/csrc -B >/dev/null 2>&1
