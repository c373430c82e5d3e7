cd $GRAFT_REPO_ROOT
WL=1080 bash tools/variants.sh "-DKF2_QUAD=0" "-DKF2_QUAD=1" 2>&1 | tail -8
make -s -C turbo_metrics_b200/csrc -B NVCCFLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC --expt-relaxed-constexpr -DKF2_QUAD=1" >/dev/null 2>&1
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bit_identical or match_oracle" 2>&1 | tail -3
