cd $GRAFT_REPO_ROOT
WL=1080 bash tools/variants.sh "-DKF2_MINB=3" "-DKF2_MINB=4" 2>&1 | tail -8
make -s -C turbo_metrics_b200/csrc -B NVCCFLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC --expt-relaxed-constexpr -DKF2_MINB=3" >/dev/null 2>&1
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bit_identical or match_oracle" 2>&1 | tail -3
ncu --set full --clock-control none --import-source on -k regex:k_frontend2 -s 2 -c 1 -o gpurun_out/r2_prof4 -f python tools/quick_time.py 3840 2160 16 16 1 64 16 > gpurun_out/r2_p4.log 2>&1
