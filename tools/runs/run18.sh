cd $GRAFT_REPO_ROOT
timeout 900 compute-sanitizer --tool racecheck --print-limit 2000 python tools/sanity_small.py > gpurun_out/race_full.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --print-limit 2000 python tools/sanity_small.py lite > gpurun_out/race_lite.log 2>&1
for f in full lite; do echo == $f; grep -c "hazard" gpurun_out/race_$f.log; grep -E "Write access at|Read access at" gpurun_out/race_$f.log | sed 's/\[.*//' | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | sort -rn | head -12; grep "RACECHECK SUMMARY" gpurun_out/race_$f.log; done
