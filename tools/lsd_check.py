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
print("frames ready", flush=True)
ref = [pyoracle.lsd_detect(fr[i]) for i in range(a.frames)]
print("oracle ready", flush=True)
for rep in range(a.repeat):
    for f0 in range(0, a.frames, a.batch):
        nb = min(a.batch, a.frames - f0)
        ex.extract_batch(fr[f0:f0 + nb])
        print(f"rep {rep} batch {f0} done", flush=True)
        if True:
            import ctypes as C
            dbg = np.zeros((nb, 16), np.int32)
            api.N.check(api.N.lib().pl_line_debug_read(ex._h, C.c_int(nb), api.N.ptr(dbg)))
            for j in range(nb):
                d = dbg[j]
                if d[0] == 2:
                    st = int(d[6]) & 0xffffffff
                    cur = int(d[15]) & 0xffffffff
                    print(f"  batch overlap frame {f0 + j}: ticket {d[1]} (head {d[4]}, batch {d[7]}) seed {d[2]} n {d[3]} point #{d[5]} {d[14]:#x} stamp att {st >> 24} tk {st & 0xffffff}; pixel's stamp now att {cur >> 24} "
                          f"tk {cur & 0xffffff}; poison {d[8]:#x} thief {d[9]} deps {[(int(x) & 0xffffffff) >> 24 for x in d[10:14]]} {[(int(x) & 0xffffff) - 1 for x in d[10:14]]}")
                elif d[0]:
                    st = int(d[6]) & 0xffffffff
                    print(f"  shadow mismatch frame {f0 + j}: ticket {d[1]} seed {d[2]} spec n {d[3]} seq n {d[4]} first diff {d[5]} stamp att {st >> 24} tk {st & 0xffffff} n0 {d[7]} "
                          f"poison {d[8]:#x} thief {d[9]} deps {[(int(x) & 0xffffffff) >> 24 for x in d[10:14]]} {[(int(x) & 0xffffff) - 1 for x in d[10:14]]} pts {d[14]:#x} {d[15]:#x}")
        for j in range(nb):
            lines, width, prec, nfa = ex.lsd_segments(frame=j)
            ol, ow, op, on = ref[f0 + j]
            same = lines.shape == ol.shape and np.abs(lines - ol).max() <= 1e-3 and np.abs(width - ow).max() <= 1e-6 and np.array_equal(prec, op)
            if not same and (int(os.environ.get("PLSLAM_LSD_DEBUG", "0")) & 256):
                import ctypes as C
                log = np.zeros((40000, 4), np.int32)
                nlog = C.c_int()
                api.N.check(api.N.lib().pl_line_debug_log(ex._h, C.c_int(j), api.N.ptr(log), C.c_int(40000), C.byref(nlog)))
                log = log[:nlog.value]
                tr = pyoracle.lsd_trace(fr[f0 + j])
                m = min(len(log), len(tr))
                neq = (log[:m, 0] != tr[:m, 0]) | ((log[:m, 1] != tr[:m, 1]) & (log[:m, 1] >= 0)) | (log[:m, 2] != tr[:m, 2])
                if neq.any() or len(log) != len(tr):
                    k = int(np.argmax(neq)) if neq.any() else m
                    print(f"  trace: {len(log)} regions vs oracle {len(tr)}; first divergence at region {k}")
                    for q in range(max(0, k - 2), min(m, k + 3)):
                        st = int(log[q, 3]) & 0xffffffff
                        print(f"    #{q}: gpu seed {log[q,0]} first {log[q,1]} final {log[q,2]} att {(st >> 24) & 63} tk {(st & 0xffffff) - 1} batch {st >> 31} poisoned {(st >> 30) & 1} | oracle seed {tr[q,0]} first {tr[q,1]} final {tr[q,2]} rect {tr[q,3]}")
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
