cd $GRAFT_REPO_ROOT
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_b2_4k.json 2> gpurun_out/r2_b2.err
tail -c 600 gpurun_out/r2_b2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_b2_4k.json'))
print('value',d['value'],'e2e',d['e2e']['value'],d['e2e']['h2d_gbs'],d['e2e']['h2d_gbs_raw'],'frac',d['roofline']['frac'])
print('parity',d['parity'])
for k,v in d['roofline']['kernels'].items(): print(k,{a:b for a,b in v.items() if a!='bound'})
for k,v in d['workloads'].items(): print(k,v['value'],v['roofline_frac'],v['e2e']['value'],v['parity'] and (v['parity']['dscore'],v['parity'].get('max_rel_norm')),v['kernel_ms_per_launch'])
print(d['clocks'], d['config']['numa'])
PY
ncu --set full --clock-control none --import-source on -k regex:k_ -s 6 -c 3 -o gpurun_out/r2_prof7 -f python tools/quick_time.py 3840 2160 16 16 1 64 16 > gpurun_out/r2_p7.log 2>&1
ncu --set full --clock-control none -k regex:k_ -s 6 -c 3 -o gpurun_out/r2_prof7_1080 -f python tools/quick_time.py 1920 1080 8 32 1 128 32 > gpurun_out/r2_p7b.log 2>&1
