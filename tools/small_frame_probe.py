"""debug aid: one short-track batch through reset / update / estimate with progress lines"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
t0 = time.time()
def say(*a):
    print("[%.1fs]" % (time.time() - t0), *a, flush=True)
import mkfbodytracker_pdaf_b200 as mk
say("imported")
T, N = int(sys.argv[1]), int(sys.argv[2])
mode = sys.argv[3] if len(sys.argv) > 3 else "reset"
m = mk.Model.load(mk.LEFT_ARM_MODEL, mk.RIGHT_ARM_MODEL)
b = mk.TrackBatch(m, T, N)
say("batch")
rng = np.random.default_rng(1)
if mode == "reset":
    b.reset(rng.random(T))
else:
    b.reset(rng.random(T))
    d = b.download()
    b.upload(d["x"], d["P"])
b.sync()
say("state ready")
for fr in range(3):
    meas = np.tile(np.array([320.0, 120.0, 200.0, 300.0, 320.0, 200.0]), (T, 1)) + rng.standard_normal((T, 6))
    b.update(meas, rng.random(T), rng.random(T))
    say("update issued", fr)
    b.sync()
    say("update done", fr)
    xb, pose = b.estimate()
    say("estimate", fr, float(pose[0, 0]))
    d = b.download()
    say("download", fr, d["parents"][0].tolist(), d["status"][:8].tolist())
