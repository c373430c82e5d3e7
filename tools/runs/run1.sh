set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2_t1.log
tail -5 gpurun_out/r2_t1.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_b1_4k.json 2> gpurun_out/r2_b1.err
tail -c 1500 gpurun_out/r2_b1_4k.json
ncu --set full --clock-control none --import-source on -k regex:k_ -s 6 -c 3 -o gpurun_out/r2_prof1 -f python tools/quick_time.py 3840 2160 16 16 1 64 16 > gpurun_out/r2_p1.log 2>&1
tail -3 gpurun_out/r2_p1.log
