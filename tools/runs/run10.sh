cd $GRAFT_REPO_ROOT
WL=1080 bash tools/variants.sh "-DKF2_SERIAL=1" "-DKF2_SERIAL=0" 2>&1 | tail -8
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "libm or ieee or bit_identical or match_oracle" 2>&1 | tail -3
ncu --set full --clock-control none --import-source on -k regex:k_frontend2 -s 2 -c 1 -o gpurun_out/r2_prof6 -f python tools/quick_time.py 3840 2160 16 16 1 64 16 > gpurun_out/r2_p6.log 2>&1
