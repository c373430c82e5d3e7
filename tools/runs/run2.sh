set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bit_identical or match_oracle or matches_oracle or pipelines_agree or race_free" 2>&1 | tail -15
QUICK=1 bash tools/variants.sh "-DKF2_MINB=4" "-DKF2_MINB=3" "-DKF2_MINB=2" 2>&1 | tail -12
timeout 120 python tools/quick_time.py 1920 1080 8 32 1 128 32 | grep "rep 1"
