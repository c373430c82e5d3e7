cd $GRAFT_REPO_ROOT
nvidia-smi topo -m 2>&1 | head -12
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "shard" 2>&1 | tail -4
python bench.py --gpus 2 --steps 5 --warmup 3 --no-refgpu --no-workloads > gpurun_out/r2_b4_2gpu.json 2> gpurun_out/r2_b4.err
tail -c 300 gpurun_out/r2_b4.err
python bench.py --gpus 2 --steps 5 --warmup 3 --shard-api > gpurun_out/r2_b4_2gpu_shard.json 2> gpurun_out/r2_b4s.err
tail -c 300 gpurun_out/r2_b4s.err
python - <<'PY'
import json
for f in ('gpurun_out/r2_b4_2gpu.json','gpurun_out/r2_b4_2gpu_shard.json'):
    d=json.load(open(f))
    print(f, 'value',d['value'],'e2e',d['e2e'],'frac',d['roofline']['frac'], d['config'].get('numa'))
PY
