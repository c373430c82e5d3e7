"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): a few odd-sized pairs through the default pipeline."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import turbo_metrics_b200 as tm
from turbo_metrics_b200 import synth
lite = len(sys.argv) > 1 and sys.argv[1] == "lite"     # score-only mode (its own warp-role map in k_hv)
deep = False
for (w, h, kind) in [(203, 131, "srgb8"), (320, 180, "nv12"), (130, 70, "p016"), (256, 160, "p016"), (256, 160, "p016_deep"),
                     (200, 136, "srgb16"), (200, 136, "srgbf32"), (200, 136, "linear"), (68, 40, "linear")]:
    deep = kind == "p016_deep"
    if kind == "srgb8":
        r, d = synth.make_pair_srgb8(w, h, frame=1, seed=3)
        mk, fmt = tm.DeviceFrame.packed, tm.PixelFormat.SRGB8
    elif kind in ("srgb16", "srgbf32", "linear"):
        r8, d8 = synth.make_pair_srgb8(w, h, frame=1, seed=3)
        mk = tm.DeviceFrame.packed
        if kind == "srgb16":
            r, d, fmt = (r8.to(torch.int32) * 257).to(torch.int16), (d8.to(torch.int32) * 257).to(torch.int16), tm.PixelFormat.SRGB16
        elif kind == "srgbf32":
            r, d, fmt = r8.float() / 255, d8.float() / 255, tm.PixelFormat.SRGBF32
        else:
            r, d, fmt = (r8.float() / 255) ** 2.2, (d8.float() / 255) ** 2.2, tm.PixelFormat.LINEARF32
            r[3, 5, 1] = -1.0   # one region through the fall-back
    else:
        bits = 8 if kind == "nv12" else 16
        r, d, pitch, ch = synth.make_pair_yuv420(w, h, bits, frame=1, seed=3)
        mk = lambda t: tm.DeviceFrame.yuv420(t, pitch, ch)
        fmt = tm.PixelFormat.NV12 if bits == 8 else tm.PixelFormat.P016
    rg, dg = r.cuda(), d.cuda()
    with tm.Ssimulacra2(w, h, fmt, batch=3, ring=2, score_only=lite, input_group=2, p016_deep=deep) as m:
        ts = [m.compute(mk(rg), mk(dg)) for _ in range(5)]
        print(w, h, kind, [round(m.get_score(t), 6) for t in ts])
