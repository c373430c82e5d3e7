cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2_b6_4gpu.json 2> gpurun_out/r2_b6.err
tail -c 200 gpurun_out/r2_b6.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29552 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > gpurun_out/r2_b6_4gpu_ref.json 2>> gpurun_out/r2_b6.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_b6_4gpu.json'))
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),round(d['e2e']['h2d_gbs'],1),round(d['e2e']['h2d_gbs_raw'],1),'frac',round(d['roofline']['frac'],4),d['clocks'])
for k,v in d['workloads'].items(): print('  ',k,round(v['value'],1),round(v['roofline_frac'],3),round(v['e2e']['value'],1))
print(json.load(open('gpurun_out/r2_b6_4gpu_ref.json'))['value'])
PY
