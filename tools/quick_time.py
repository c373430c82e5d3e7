"""Quick device timing of one configuration (development aid, not the bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import turbo_metrics_b200 as tm
if os.environ.get('SSIMU2_SO'):   # development aid: time another build of the library
    import turbo_metrics_b200._lib as _l
    _l.SO_PATH = os.path.abspath(os.environ['SSIMU2_SO'])
from turbo_metrics_b200 import synth

w, h, bits = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
batch, ring, npairs, ndistinct = int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]), int(sys.argv[7])
fmt = tm.PixelFormat.NV12 if bits == 8 else tm.PixelFormat.P016
frames = []
for i in range(ndistinct):
    rb, db, pitch, ch = synth.make_pair_yuv420(w, h, bits, frame=i, seed=1, device="cuda")
    frames.append((rb, db))
F = lambda t: tm.DeviceFrame.yuv420(t, pitch, ch)
m = tm.Ssimulacra2(w, h, fmt, batch=batch, ring=ring)
info = m.info()
print("mem MB", m.mem_usage() / 2**20, "alg bytes", info.alg_bytes_per_pair)
for rep in range(2):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    ts = [m.compute(F(frames[i % ndistinct][0]), F(frames[i % ndistinct][1])) for i in range(npairs)]
    m.flush()
    scores = [m.get_score(t) for t in ts]
    e1.record()
    torch.cuda.synchronize()
    dt = time.time() - t0
    print(f"rep {rep}: {npairs / dt:.1f} pairs/s wall, {dt * 1000 / npairs:.3f} ms/pair; kernel ms/batch {m.last_batch_ms()}")
    print("  roofline frac of 6545 GB/s:", info.alg_bytes_per_pair * npairs / dt / 6545e9)
print(scores[:4])
