"""Two (or more) host threads, each with its own handles on the same GPU, creating / scoring / destroying at the same time
("different handles may be driven from different threads", include/ssimu2_b200.h).  usage: python tools/soak_threads.py [seconds] [threads]"""
import sys, os, time, random, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import turbo_metrics_b200 as tm
from turbo_metrics_b200 import synth
T = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
NT = int(sys.argv[2]) if len(sys.argv) > 2 else 3
P = tm.PixelFormat
cases = []
for (w, h, kind) in [(1920, 1080, "nv12"), (640, 360, "p016"), (512, 512, "srgb8"), (256, 254, "nv12"), (1280, 720, "srgb8")]:
    if kind == "srgb8":
        fr = [synth.make_pair_srgb8(w, h, frame=i, seed=3, device="cuda") for i in range(2)]
        mk, fmt = tm.DeviceFrame.packed, P.SRGB8
    else:
        bits = 8 if kind == "nv12" else 16
        raw = [synth.make_pair_yuv420(w, h, bits, frame=i, seed=3, device="cuda") for i in range(2)]
        pitch, ch = raw[0][2], raw[0][3]
        fr = [(a, b) for a, b, _, _ in raw]
        mk, fmt = (lambda t, p=pitch, c=ch: tm.DeviceFrame.yuv420(t, p, c)), (P.NV12 if bits == 8 else P.P016)
    with tm.Ssimulacra2(w, h, fmt, batch=1, ring=1) as m:
        want = [m.compute_sync(mk(a), mk(b)) for a, b in fr]
    cases.append((w, h, fmt, mk, fr, want))
torch.cuda.synchronize()
errors, counts = [], [0] * NT
def worker(k):
    rnd = random.Random(100 + k)
    t0 = time.time()
    try:
        while time.time() - t0 < T and not errors:
            w, h, fmt, mk, fr, want = rnd.choice(cases)
            batch, ring, so = rnd.choice([1, 4, 16, 64]), rnd.choice([1, 2, 3]), rnd.random() < 0.5
            n = rnd.randint(1, 5 * batch)
            with tm.Ssimulacra2(w, h, fmt, batch=batch, ring=ring, score_only=so) as m:
                ts = [m.compute(mk(fr[i % 2][0]), mk(fr[i % 2][1])) for i in range(n)]
                for i, t in enumerate(ts):
                    s = m.get_score(t)
                    if s != want[i % 2]:
                        errors.append((k, w, h, batch, ring, so, i, s, want[i % 2])); return
            counts[k] += n
    except Exception as e:
        errors.append((k, repr(e)))
th = [threading.Thread(target=worker, args=(k,)) for k in range(NT)]
[t.start() for t in th]; [t.join() for t in th]
print("errors", errors[:3]) if errors else print(f"threads ok: {NT} threads, {sum(counts)} pairs, {T:.0f} s")
sys.exit(1 if errors else 0)
