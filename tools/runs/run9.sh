cd $GRAFT_REPO_ROOT
WL=1080 bash tools/variants.sh "-DKF2_THREADS=128 -DKF2_MINB=4 -DKF2_SERIAL=1" "-DKF2_THREADS=128 -DKF2_MINB=4 -DKF2_SERIAL=0" "-DKF2_THREADS=128 -DKF2_MINB=5 -DKF2_SERIAL=1" "-DKF2_THREADS=256 -DKF2_MINB=3 -DKF2_SERIAL=1" 2>&1 | tail -14
make -s -C turbo_metrics_b200/csrc -B NVCCFLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC --expt-relaxed-constexpr -DKF2_THREADS=128 -DKF2_MINB=4" >/dev/null 2>&1
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "libm or ieee or bit_identical or match_oracle" 2>&1 | tail -3
ncu --set full --clock-control none --import-source on -k regex:k_frontend2 -s 2 -c 1 -o gpurun_out/r2_prof5 -f python tools/quick_time.py 3840 2160 16 16 1 64 16 > gpurun_out/r2_p5.log 2>&1
