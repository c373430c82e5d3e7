"""Soak: random (format, size, batch, ring, mode) configurations for a fixed time; every repetition of a pair must have the bits
of the score that a batch-1 / ring-1 full-mode handle gives for it.  A race in the strip hand-off, the ring, the ticket
bookkeeping or the lite role map shows up as a score that depends on timing.  usage: python tools/soak.py [seconds]"""
import sys, os, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import turbo_metrics_b200 as tm
from turbo_metrics_b200 import synth
T = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rnd = random.Random(7)
P = tm.PixelFormat
sizes = [(3840, 2160), (1920, 1080), (1280, 720), (512, 512), (640, 360), (257, 255), (2560, 1440), (720, 480)]
if os.environ.get("SOAK_BIG"):      # only the large geometries (many strips x many bands per chain)
    sizes = [(3840, 2160), (1920, 1080), (2560, 1440)]
t0 = time.time(); runs = 0; pairs = 0
while time.time() - t0 < T:
    w, h = rnd.choice(sizes)
    kind = rnd.choice(["nv12", "p016", "p016_12", "srgb8", "linear", "srgb16", "srgbf32"])
    deep = False
    nd = 3
    if kind in ("nv12", "p016", "p016_12"):
        bits = 8 if kind == "nv12" else 16
        w2, h2 = w & ~1, h & ~1
        fr = [synth.make_pair_yuv420(w2, h2, bits, frame=i, seed=runs + 1, device="cuda") for i in range(nd)]
        pitch, ch = fr[0][2], fr[0][3]
        mk, fmt, w, h = (lambda t: tm.DeviceFrame.yuv420(t, pitch, ch)), (P.NV12 if bits == 8 else P.P016), w2, h2
        fr = [(a, b) for a, b, _, _ in fr]
        if kind == "p016_12":     # 12 significant bits; with or without the hint
            for a, b in fr:
                a.view(torch.int16).__ior__(torch.randint(0, 4, a.view(torch.int16).shape, device="cuda", dtype=torch.int16) << 4)
                b.view(torch.int16).__ior__(torch.randint(0, 4, b.view(torch.int16).shape, device="cuda", dtype=torch.int16) << 4)
            deep = rnd.random() < 0.7
    else:
        base = [synth.make_pair_srgb8(w, h, frame=i, seed=runs + 1, device="cuda") for i in range(nd)]
        mk = tm.DeviceFrame.packed
        if kind == "srgb8":
            fr, fmt = base, P.SRGB8
        elif kind == "srgb16":
            fr, fmt = [((a.to(torch.int32) * 257).to(torch.int16), (b.to(torch.int32) * 257).to(torch.int16)) for a, b in base], P.SRGB16
        elif kind == "srgbf32":
            fr, fmt = [(a.float() / 255, b.float() / 255) for a, b in base], P.SRGBF32
        else:
            fr, fmt = [((a.float() / 255) ** 2.2, (b.float() / 255) ** 2.2) for a, b in base], P.LINEARF32
    with tm.Ssimulacra2(w, h, fmt, batch=1, ring=1) as m:
        want = [m.compute_sync(mk(a), mk(b)) for a, b in fr]
    big = w * h > 3_000_000
    batch = rnd.choice([1, 2, 5, 8, 16] if big else [1, 3, 8, 16, 32, 64])
    ring = rnd.choice([1, 2, 3, 4])
    so = rnd.random() < 0.5
    grp = rnd.choice([0, 0, 1, 3])
    n = rnd.randint(batch, 6 * batch) if big else rnd.randint(batch, 12 * batch)
    host = rnd.random() < 0.25          # a quarter of the runs through ssimu2_submit_host (pinned host copies of the frames)
    hfr = [(a.cpu().pin_memory(), b.cpu().pin_memory()) for a, b in fr] if host else None
    print(f"run {runs}: {kind} {w}x{h} batch {batch} ring {ring} score_only {so} input_group {grp} n {n} deep {deep} host {host}", flush=True)
    with tm.Ssimulacra2(w, h, fmt, batch=batch, ring=ring, score_only=so, input_group=grp, p016_deep=deep) as m:
        ts = []
        got = {}
        for i in range(n):
            a, b = (hfr if host else fr)[i % nd]
            ts.append(m.compute_from_cpu(mk(a), mk(b)) if host else m.compute(mk(a), mk(b)))
            if rnd.random() < 0.1:      # interleaved fetches of an older ticket
                j = rnd.randrange(len(ts))
                got[j] = m.get_score(ts[j])
        for j, t in enumerate(ts):
            s = got[j] if j in got else m.get_score(t)
            assert s == want[j % nd], (kind, w, h, batch, ring, so, grp, j, s, want[j % nd])
    runs += 1; pairs += n
print(f"soak ok: {runs} configurations, {pairs} pairs, {time.time() - t0:.0f} s")
