cd $GRAFT_REPO_ROOT
nvidia-smi topo -m > gpurun_out/r2_topo8.txt 2>&1
(lscpu | grep -iE "numa|socket|model name|^CPU\(s\)"; cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c) >> gpurun_out/r2_topo8.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "shard" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 --no-refgpu --no-workloads > gpurun_out/r2_b5_8gpu.json 2> gpurun_out/r2_b5.err
tail -c 300 gpurun_out/r2_b5.err
python bench.py --gpus 8 --steps 5 --warmup 3 --shard-api > gpurun_out/r2_b5_8gpu_shard.json 2> gpurun_out/r2_b5s.err
tail -c 300 gpurun_out/r2_b5s.err
python - <<'PY'
import json
for f in ('gpurun_out/r2_b5_8gpu.json','gpurun_out/r2_b5_8gpu_shard.json'):
    try:
        d=json.load(open(f))
        print(f, 'value',round(d['value'],1),'e2e',{k:(round(v,1) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k!='note'},'frac',round(d['roofline']['frac'],4), d['config'].get('numa'), d.get('clocks'))
    except Exception as e: print(f, 'ERR', e)
PY
