"""Times baseline/refgpu (the reference's GPU design) on synthetic pairs: usage refgpu_time.py W H BITS NPAIRS NDISTINCT"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from turbo_metrics_b200 import synth
from baseline.refgpu import refgpu

w, h, bits, n, nd = (int(x) for x in sys.argv[1:6])
fr = [synth.make_pair_yuv420(w, h, bits, frame=i, seed=1, device="cuda") for i in range(nd)]
pitch, ch = fr[0][2], fr[0][3]
with refgpu.RefGpu(w, h, bits) as r:
    print(r.info())
    for i in range(4):
        r.compute(fr[i % nd][0], fr[i % nd][1], pitch, ch)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sc = [r.compute(fr[i % nd][0], fr[i % nd][1], pitch, ch)[0] for i in range(n)]
    dt = time.perf_counter() - t0
    print(f"{n / dt:.1f} pairs/s, {dt * 1000 / n:.3f} ms/pair; scores {sc[:3]}")
