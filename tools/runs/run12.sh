cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "score_only" 2>&1 | tail -6
python bench.py --steps 5 --warmup 3 --score-only --no-refgpu > gpurun_out/r2_b3_so.json 2> gpurun_out/r2_b3.err
tail -c 400 gpurun_out/r2_b3.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_b3_so.json'))
print('value',d['value'],'frac',d['roofline']['frac'],d['parity'])
for k,v in d['roofline']['kernels'].items(): print(k,v['ms_per_launch'])
for k,v in d['workloads'].items(): print(k,v['value'],v['roofline_frac'],v['kernel_ms_per_launch'])
PY
