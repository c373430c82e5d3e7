#!/bin/bash
# usage: tools/variants_bench.sh "<-D flags variant 1>" ...   (GPU box; rebuilds and runs the sustained bench.py loop for each)
cd "$(dirname "$0")/.."
for v in "$@"; do
  make -s -C turbo_metrics_b200/csrc -B NVCCFLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC --expt-relaxed-constexpr $v" >/dev/null 2>&1 || { echo "BUILD FAILED: $v"; continue; }
  python bench.py --no-cpu-baseline --no-refgpu --steps 8 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('== $v', round(d['value'],1), round(d['roofline']['pipeline']['frac'],4), d['roofline']['kernel_ms_per_launch'], d['clocks'])"
done
make -s -C turbo_metrics_b200/csrc -B >/dev/null 2>&1
