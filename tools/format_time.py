"""Front-end / k_hv device time per launch for every pixel format at one size (development aid).
usage: python tools/format_time.py [w h batch]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import turbo_metrics_b200 as tm
if os.environ.get('SSIMU2_SO'):
    import turbo_metrics_b200._lib as _l
    _l.SO_PATH = os.path.abspath(os.environ['SSIMU2_SO'])
from turbo_metrics_b200 import synth
w, h, batch = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080, 32)
P = tm.PixelFormat
for kind in ("nv12", "p016", "p016_12bit", "srgb8", "srgb16", "srgbf32", "linear"):
    if kind in ("nv12", "p016", "p016_12bit"):
        bits = 8 if kind == "nv12" else 16
        r, d, pitch, ch = synth.make_pair_yuv420(w, h, bits, frame=1, seed=3, device="cuda")
        if kind == "p016_12bit":   # 12 significant bits in the 16-bit containers (HEVC Main12 through NVDEC): no exact R / B memo
            g = torch.Generator(device="cuda").manual_seed(1)
            for t in (r, d):
                v = t.view(torch.int16)
                v |= (torch.randint(0, 4, v.shape, generator=g, device="cuda", dtype=torch.int16) << 4)
        mk, fmt = (lambda t: tm.DeviceFrame.yuv420(t, pitch, ch)), (P.NV12 if bits == 8 else P.P016)
    else:
        r8, d8 = synth.make_pair_srgb8(w, h, frame=1, seed=3, device="cuda")
        mk = tm.DeviceFrame.packed
        if kind == "srgb8":
            r, d, fmt = r8, d8, P.SRGB8
        elif kind == "srgb16":
            r, d, fmt = (r8.to(torch.int32) * 257).to(torch.int16), (d8.to(torch.int32) * 257).to(torch.int16), P.SRGB16
        elif kind == "srgbf32":
            r, d, fmt = r8.float() / 255, d8.float() / 255, P.SRGBF32
        else:
            r, d, fmt = (r8.float() / 255) ** 2.2, (d8.float() / 255) ** 2.2, P.LINEARF32
    for deep in ((False, True) if kind == "p016_12bit" else (False,)):
      with tm.Ssimulacra2(w, h, fmt, batch=batch, ring=1, p016_deep=deep) as m:
        for rep in range(2):
            ts = m.compute_batch([mk(r)] * (2 * batch), [mk(d)] * (2 * batch))
            sc = m.get_scores(ts)
        ms = m.last_batch_ms()
      print(f"{kind + (' +P016_DEEP' if deep else ''):22s} {w}x{h} batch {batch}: front-end {ms[0]:.3f} ms, k_hv {ms[1]:.3f} ms, score {sc[0]:.6f}", flush=True)
