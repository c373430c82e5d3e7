cd $GRAFT_REPO_ROOT
for b in 16 24 32 48; do
python bench.py --steps 5 --warmup 3 --batch $b --no-workloads --no-cpu-baseline --no-refgpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']; print('batch', $b, 'value', round(d['value'],1), 'frac', round(d['roofline']['frac'],4), {a:round(v['ms_per_launch'],3) for a,v in k.items()}, d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
