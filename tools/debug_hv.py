"""Development aid: norms of the fused H+V pipeline against the split pipeline, per scale / channel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import turbo_metrics_b200 as tm
from turbo_metrics_b200 import synth
sizes = [(64, 24), (64, 64), (128, 24), (200, 100), (512, 512)]
for w, h in sizes:
    r, d = synth.make_pair_srgb8(w, h, frame=1, seed=3)
    rg, dg = r.cuda(), d.cuda()
    res = {}
    for pl in ("split", "hv"):
        with tm.Ssimulacra2(w, h, tm.PixelFormat.SRGB8, batch=1, ring=1, pipeline=pl) as m:
            t = m.compute(tm.DeviceFrame.packed(rg), tm.DeviceFrame.packed(dg))
            res[pl] = (m.get_score(t), m.get_norms(t), m.info().nscales)
    a, b = res["split"][1], res["hv"][1]
    ns = res["hv"][2]
    print(f"== {w}x{h}: score split {res['split'][0]:.6f} hv {res['hv'][0]:.6f} nscales {ns}")
    for s in range(ns):
        for c in range(3):
            idx = [c * 36 + s * 6 + k for k in range(6)]
            rel = np.abs(a[idx] - b[idx]) / np.maximum(np.abs(a[idx]), 1e-30)
            if rel.max() > 1e-7:
                print(f"   scale {s} ch {c}: rel err [L1 ssim, art, det | L4 ssim, art, det] = " + " ".join(f"{x:.2e}" for x in rel))
