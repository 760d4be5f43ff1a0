"""Segment-set equality sweep: LSD segments of the CUDA path vs the CPU oracle, frame by frame, for a grower configuration
taken from the environment (PLSLAM_LSD_*).  Prints the frames that differ."""
import argparse
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
pkg = importlib.import_module("orb_slam2_modification_with-point-and-line-feature_b200")
api = pkg.load_api()
import pyoracle

ap = argparse.ArgumentParser()
ap.add_argument("--seed", type=int, default=4242)
ap.add_argument("--frames", type=int, default=5)
ap.add_argument("--batch", type=int, default=5)
ap.add_argument("--w", type=int, default=640)
ap.add_argument("--h", type=int, default=480)
ap.add_argument("--repeat", type=int, default=1)
a = ap.parse_args()
fr = pkg.synth.frames(a.seed, a.frames, a.w, a.h)
ex = api.LineExtractor(max_cols=a.w, max_rows=a.h, max_batch=a.batch)
bad = 0
ref = [pyoracle.lsd_detect(fr[i]) for i in range(a.frames)]
for rep in range(a.repeat):
    for f0 in range(0, a.frames, a.batch):
        nb = min(a.batch, a.frames - f0)
        ex.extract_batch(fr[f0:f0 + nb])
        for j in range(nb):
            lines, width, prec, nfa = ex.lsd_segments(frame=j)
            ol, ow, op, on = ref[f0 + j]
            same = lines.shape == ol.shape and np.abs(lines - ol).max() <= 1e-3 and np.abs(width - ow).max() <= 1e-6 and np.array_equal(prec, op)
            if not same:
                bad += 1
                first = -1
                if lines.shape == ol.shape:
                    first = int(np.argmax(np.abs(lines - ol).max(axis=1) > 1e-3))
                else:
                    m = min(len(lines), len(ol))
                    d = np.abs(lines[:m] - ol[:m]).max(axis=1) > 1e-3
                    first = int(np.argmax(d)) if d.any() else m
                nd = int((np.abs(lines - ol).max(axis=1) > 1e-3).sum()) if lines.shape == ol.shape else -1
                print(f"rep {rep} frame {f0 + j}: {lines.shape[0]} segments vs oracle {ol.shape[0]}, first difference at segment {first}, {nd} segments differ")
print(f"{a.frames * a.repeat} frame extractions, {bad} differ from the oracle")
