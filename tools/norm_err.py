"""Largest relative deviation of the 108 norms (and the score) from the CPU oracle on a few cases (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import turbo_metrics_b200 as tm
from turbo_metrics_b200 import synth
from oracle import oracle
oracle.build()
for (w, h, bits) in [(640, 360, 8), (960, 540, 16), (1920, 1080, 8)]:
    rb, db, pitch, ch = synth.make_pair_yuv420(w, h, bits, frame=2, seed=7)
    so, no, _ = oracle.ssimu2_yuv420(rb.numpy(), db.numpy(), pitch, ch, w, h, bits)
    fmt = tm.PixelFormat.NV12 if bits == 8 else tm.PixelFormat.P016
    with tm.Ssimulacra2(w, h, fmt, batch=2, ring=1) as m:
        t = m.compute(tm.DeviceFrame.yuv420(rb.cuda(), pitch, ch), tm.DeviceFrame.yuv420(db.cuda(), pitch, ch))
        s, n = m.get_score(t), m.get_norms(t)
    rel = np.abs(n - no) / np.maximum(np.abs(no), 1e-300)
    print(f"{w}x{h} {bits}-bit: max rel norm err {rel.max():.2e}, score err {abs(s - so):.2e}")
