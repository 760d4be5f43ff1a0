"""Where the end-to-end step (host buffers through the public API + caller glue) spends its time.  Run on a GPU box."""
import cProfile
import importlib
import os
import pstats
import sys
import time

import numpy as np

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
tfe = fe.TrackingFrontEnd(gb)


def step():
    t0 = time.perf_counter()
    a = gb.extract_orb(gray)
    t1 = time.perf_counter()
    b = gb.extract_lines(gray)
    t2 = time.perf_counter()
    s = tfe.run(gray, depth, Tcw, sf, features=(a, b))
    t3 = time.perf_counter()
    return (t1 - t0, t2 - t1, t3 - t2), s


step()
for _ in range(2):
    print("orb %.1f ms  lines %.1f ms  track(run) %.1f ms" % tuple(1e3 * x for x in step()[0]))
pr = cProfile.Profile()
pr.enable()
step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)

import concurrent.futures
pool = concurrent.futures.ThreadPoolExecutor(1)
for _ in range(4):
    t0 = time.perf_counter()
    a = gb.extract_orb(gray)
    fl = pool.submit(gb.extract_lines, gray)
    s = tfe.run(gray, depth, Tcw, sf, features=(a, fl))
    print("pipelined step (ORB, then lines in a second thread || point-side glue) %.1f ms" % (1e3 * (time.perf_counter() - t0)))
