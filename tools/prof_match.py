"""Times the four batched matcher calls of the 300-frame schedule (C3, C2, D3, D5) separately.  Run on a GPU box."""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "orb_slam2_modification_with-point-and-line-feature_b200"
pkg = importlib.import_module(PKG)
api = pkg.load_api()
fe = importlib.import_module(PKG + ".frontend")
F = int(sys.argv[1]) if len(sys.argv) > 1 else 300
gray, depth, Tcw = pkg.synth.room_sequence(F, 640, 480, workers=min(32, os.cpu_count() or 1))
gb = fe.GpuBackend(api, 480, 640, 1000, chunk=F, device=0)
sf = gb.scale_factors()
calls = []


class Rec:
    def __init__(self, gb):
        self.gb = gb

    def __getattr__(self, k):
        f = getattr(self.gb, k)
        if k.endswith("_batch"):
            def w(*a):
                calls.append((k, a))
                return f(*a)
            return w
        return f


tfe = fe.TrackingFrontEnd(Rec(gb))
feats = (gb.extract_orb(gray), gb.extract_lines(gray))
tfe.run(gray, depth, Tcw, sf, features=feats)
for rep in range(3):
    out = []
    for k, a in calls:
        gb.m.sync()
        t0 = time.perf_counter()
        getattr(gb, k)(*a)
        gb.m.sync()
        out.append("%s %.2f ms (%d launches)" % (k, 1e3 * (time.perf_counter() - t0), gb.m.last_launches()))
    print(" | ".join(out))
