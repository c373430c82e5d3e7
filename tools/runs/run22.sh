cd $GRAFT_REPO_ROOT
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_frontend2|k_hv|k_finalize|k_build_eotf_lut|k_hpass|k_vpass' -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-workloads --no-cpu-baseline --no-refgpu > gpurun_out/r2_ncu_bench.log 2>&1
tail -2 gpurun_out/r2_launches.csv | cut -c1-200
