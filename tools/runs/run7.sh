cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "libm or ieee or bit_identical or match_oracle" 2>&1 | tail -3
WL=1080 bash tools/variants.sh "-DKF2_THREADS=256 -DKF2_MINB=3" "-DKF2_THREADS=128 -DKF2_MINB=5" "-DKF2_THREADS=128 -DKF2_MINB=4 -DKF2_SERIAL=0" "-DKF2_THREADS=256 -DKF2_MINB=4" 2>&1 | tail -14
