#!/bin/bash
# usage: tools/variants.sh "<-D flags variant 1>" "<variant 2>" ...   (runs on the GPU box; rebuilds and times each)
# each variant is timed with ring 1 (per-kernel ms: front-end, k_hv, -, finalize) at 4K P016, 16 pairs per launch;
# WL=1080 / WL=512 add the other workloads
cd "$(dirname "$0")/.."
for v in "$@"; do
  make -s -C turbo_metrics_b200/csrc -B NVCCFLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC --expt-relaxed-constexpr $v" >/dev/null 2>&1 || { echo "BUILD FAILED: $v"; continue; }
  echo "== $v"
  timeout 120 python tools/quick_time.py 3840 2160 16 16 1 96 16 | grep "rep 1"
  [ -z "$WL" ] || timeout 120 python tools/quick_time.py 1920 1080 8 32 1 128 32 | grep "rep 1"
done
make -s -C turbo_metrics_b200/csrc -B >/dev/null 2>&1
