"""Parity against the oracle at sizes beyond the test suite: 8K P016 (7680x4320) and an odd 4098x2162 NV12 frame, full and
score-only mode (run on the GPU box; the oracle takes 30 s at 8K).  Round 2: score 1e-9, norms 3e-9 / 1.7e-8."""
import sys, os, time
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import turbo_metrics_b200 as tm
from turbo_metrics_b200 import synth
from oracle import oracle
for (w, h, bits) in [(7680, 4320, 16), (4098, 2162, 8)]:
    rb, db, pitch, ch = synth.make_pair_yuv420(w, h, bits, frame=0, seed=2)
    t0 = time.time(); so, no, _ = oracle.ssimu2_yuv420(rb.numpy(), db.numpy(), pitch, ch, w, h, bits); t1 = time.time()
    fmt = tm.PixelFormat.P016 if bits == 16 else tm.PixelFormat.NV12
    F = lambda t: tm.DeviceFrame.yuv420(t, pitch, ch)
    for so_mode in (False, True):
        with tm.Ssimulacra2(w, h, fmt, batch=2, ring=2, score_only=so_mode) as m:
            rg, dg = rb.cuda(), db.cuda()
            ts = [m.compute(F(rg), F(dg)) for _ in range(5)]
            sc = [m.get_score(t) for t in ts]
            rel = None
            if not so_mode:
                n = m.get_norms(ts[0]); nz = no != 0
                rel = float((np.abs(n[nz] - no[nz]) / no[nz]).max())
            print(w, h, bits, 'score_only' if so_mode else 'full', sc[0], so, abs(sc[0] - so), rel, len(set(sc)), 'mem GB', m.mem_usage() / 2**30, 'oracle s', round(t1 - t0, 1))
            assert abs(sc[0] - so) <= 0.01 and len(set(sc)) == 1 and (rel is None or rel <= 1e-4)
print('ok')
