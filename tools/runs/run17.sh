cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanity_small.py 2>&1 | tail -6
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanity_small.py lite 2>&1 | tail -6
