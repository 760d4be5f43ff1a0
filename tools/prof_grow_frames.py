"""Per-frame accounting of k_lsd_grow2 over a batch: how long each frame was active and what its sequencer did."""
import argparse
import ctypes as C
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("orb_slam2_modification_with-point-and-line-feature_b200")
api = pkg.load_api()
N = api.N
ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=300)
ap.add_argument("--room", action="store_true")
a = ap.parse_args()
w, h = 640, 480
fr = pkg.synth.room_sequence(a.frames, w, h, workers=min(32, os.cpu_count() or 1))[0] if a.room else pkg.synth.frames(6000, a.frames, w, h)
d_in = torch.from_numpy(fr).cuda()
ex = api.LineExtractor(max_cols=w, max_rows=h, max_batch=a.frames)
ML = 80
d_kls = torch.empty((a.frames, ML, 17), dtype=torch.float32, device="cuda")
d_desc = torch.empty((a.frames, ML, 32), dtype=torch.uint8, device="cuda")
d_co = torch.empty((a.frames, ML, 3), dtype=torch.float64, device="cuda")
d_n = torch.empty(a.frames, dtype=torch.int32, device="cuda")
N.check(N.lib().pl_line_set_profiling(ex._h, 1))
for _ in range(2):
    ex.extract_batch_dev(d_in.data_ptr(), a.frames, h, w, w, w * h, ML, d_kls.data_ptr(), d_desc.data_ptr(), d_co.data_ptr(), d_n.data_ptr())
    ex.sync()
rows = []
for f in range(a.frames):
    ph = np.zeros(16, np.int64)
    N.check(N.lib().pl_line_grow_phases(ex._h, C.c_int(f), N.ptr(ph)))
    rows.append(ph)
P = np.array(rows, dtype=np.float64)
act = P[:, 1] / 1e6
print("active Mcycles: min %.1f p10 %.1f p50 %.1f p90 %.1f max %.1f mean %.1f" % (act.min(), np.percentile(act, 10), np.percentile(act, 50), np.percentile(act, 90), act.max(), act.mean()))
seq = (P[:, 2] + P[:, 0] + P[:, 8]) / 1e6
print("sequencer busy (commit + regrow + issue) Mcycles: p50 %.1f max %.1f; share of active: p50 %.2f" % (np.percentile(seq, 50), seq.max(), np.percentile(seq / act, 50)))
print("tickets p50 %.0f max %.0f; corr(active, tickets) %.2f" % (np.percentile(P[:, 3], 50), P[:, 3].max(), np.corrcoef(act, P[:, 3])[0, 1]))
o = np.argsort(-act)[:8]
for f in o:
    print("frame %3d active %.1f commit %.1f regrow %.1f issue %.1f idle %.1f grow %.1f wait %.1f tickets %d regrown %d" % (f, act[f], P[f, 2] / 1e6, P[f, 0] / 1e6, P[f, 8] / 1e6, P[f, 9] / 1e6, P[f, 10] / 1e6, P[f, 11] / 1e6, P[f, 3], P[f, 4]))
print("first 8 by index:", np.round(act[:8], 1), " last 8:", np.round(act[-8:], 1))
