"""Per-frame accounting of k_lsd_grow2 over a batch (pl_line_grow_phases): distribution of the frames' active time, sequencer and
grower cycles.  usage: python tools/prof_frames.py [--frames N] [--room]"""
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
W, H = 640, 480
fr = pkg.synth.room_sequence(a.frames, W, H, workers=min(32, os.cpu_count() or 1))[0] if a.room else pkg.synth.frames(6000, a.frames, W, H)
d_in = torch.from_numpy(fr).cuda()
ex = api.LineExtractor(max_cols=W, max_rows=H, max_batch=a.frames)
ML = 80
d_kls = torch.empty((a.frames, ML, 17), dtype=torch.float32, device="cuda")
d_desc = torch.empty((a.frames, ML, 32), dtype=torch.uint8, device="cuda")
d_co = torch.empty((a.frames, ML, 3), dtype=torch.float64, device="cuda")
d_n = torch.empty(a.frames, dtype=torch.int32, device="cuda")
N.check(N.lib().pl_line_set_profiling(ex._h, 1))
for _ in range(2):
    ex.extract_batch_dev(d_in.data_ptr(), a.frames, H, W, W, W * H, ML, d_kls.data_ptr(), d_desc.data_ptr(), d_co.data_ptr(), d_n.data_ptr())
    ex.sync()
ph = np.zeros((a.frames, 16), np.int64)
for f in range(a.frames):
    N.check(N.lib().pl_line_grow_phases(ex._h, C.c_int(f), N.ptr(ph[f])))
M = 1e6
act = ph[:, 1] / M
print(f"frames {a.frames}: active Mcycles min {act.min():.1f} median {np.median(act):.1f} mean {act.mean():.1f} p90 {np.percentile(act, 90):.1f} max {act.max():.1f}")
print(f"tickets mean {ph[:, 3].mean():.0f} max {ph[:, 3].max()}; committed mean {ph[:, 5].mean():.0f}; regrown mean {ph[:, 4].mean():.0f}")
print(f"growing Mcycles (sum over growers) mean {ph[:, 10].mean() / M:.1f} max {ph[:, 10].max() / M:.1f}; waiting mean {ph[:, 11].mean() / M:.1f}; "
      f"sequencer commit {ph[:, 2].mean() / M:.1f} regrow {ph[:, 0].mean() / M:.1f} issue {ph[:, 8].mean() / M:.1f} idle {ph[:, 9].mean() / M:.1f}")
order = np.argsort(-act)[:8]
print("slowest frames:", [(int(i), round(float(act[i]), 1), int(ph[i, 3])) for i in order])
