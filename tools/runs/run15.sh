cd $GRAFT_REPO_ROOT
ncu --set full --clock-control none --import-source on -k regex:k_frontend2 -s 2 -c 1 -o gpurun_out/r2_prof8 -f python tools/quick_time.py 3840 2160 16 16 1 64 16 > gpurun_out/r2_p8.log 2>&1
tail -4 gpurun_out/r2_p8.log
