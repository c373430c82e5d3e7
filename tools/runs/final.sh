cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:k_ -s 6 -c 3 -o gpurun_out/r2_final_4k -f python tools/quick_time.py 3840 2160 16 16 1 64 16 > gpurun_out/r2_pf1.log 2>&1
ncu --set full --clock-control none -k regex:k_ -s 6 -c 3 -o gpurun_out/r2_final_1080p -f python tools/quick_time.py 1920 1080 8 32 1 128 32 > gpurun_out/r2_pf2.log 2>&1
python tools/summarize_ncu.py gpurun_out/r2_final_4k.ncu-rep r2_final_4k 4k 16 "4K P016 3840x2160, k_frontend2 + k_hv + k_finalize (round-2 final), 16 frame pairs per launch" > /dev/null
python tools/summarize_ncu.py gpurun_out/r2_final_1080p.ncu-rep r2_final_1080p 1080p 32 "1080p NV12 1920x1080 (round-2 final)" > /dev/null
cp profiles/traffic_r2.json profiles/r2_final_4k_ncu_summary.md profiles/r2_final_1080p_ncu_summary.md gpurun_out/   # profiles/ does not travel back: copy by hand
python bench.py > gpurun_out/r2_bench_4k.json 2> gpurun_out/r2_bench.err; tail -c 300 gpurun_out/r2_bench.err
python bench.py --score-only --no-refgpu > gpurun_out/r2_bench_4k_score_only.json 2>> gpurun_out/r2_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2>> gpurun_out/r2_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_frontend2|k_hv|k_finalize|k_build_eotf_lut|k_hpass|k_vpass' -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-workloads --no-cpu-baseline --no-refgpu > gpurun_out/r2_ncu_bench.log 2>&1
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench_4k.json','gpurun_out/r2_bench_4k_score_only.json','gpurun_out/r2_bench_reference.json'):
    d=json.load(open(f))
    print(f,'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',d.get('roofline',{}).get('frac'),d.get('clocks'))
    if 'workloads' in d:
        for k,v in d['workloads'].items(): print('   ',k,round(v['value'],1),round(v['roofline_frac'],3),round(v['e2e']['value'],1),v['parity'] and v['parity']['dscore'])
    if d.get('parity'): print('   parity',d['parity']['dscore'],d['parity'].get('max_rel_norm'))
PY
