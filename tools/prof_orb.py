"""Stage timing of the batched ORB path with inputs resident in HBM (CUDA events on the handle's stream)."""
import argparse
import ctypes as C
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("orb_slam2_modification_with-point-and-line-feature_b200")
api = pkg.load_api()
N = api.N

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=256)
ap.add_argument("--chunk", type=int, default=32)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--w", type=int, default=640)
ap.add_argument("--h", type=int, default=480)
ap.add_argument("--nf", type=int, default=1000)
a = ap.parse_args()

fr = pkg.synth.frames(6000, a.frames, a.w, a.h)
d_in = torch.from_numpy(fr).cuda()
ex = api.ORBextractor(a.nf, 1.2, 8, 20, 7, max_cols=a.w, max_rows=a.h, max_batch=a.chunk)
cap = ex.max_keypoints()
d_kps = torch.empty((a.frames, cap, 7), dtype=torch.float32, device="cuda")
d_desc = torch.empty((a.frames, cap, 32), dtype=torch.uint8, device="cuda")
d_n = torch.empty(a.frames, dtype=torch.int32, device="cuda")
st = torch.cuda.ExternalStream(ex.stream())
torch.cuda.synchronize()


def run():
    ex.extract_batch_dev(d_in.data_ptr(), a.frames, a.h, a.w, a.w, a.w * a.h, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_n.data_ptr())


for _ in range(3):
    run()
ex.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(a.iters):
    run()
e1.record(st)
ex.sync()
ms = e0.elapsed_time(e1) / a.iters
print(f"total: {ms:.3f} ms / {a.frames} frames = {ms*1000/a.frames:.2f} us/frame = {a.frames/ms*1000:.0f} frames/s (chunk {a.chunk})")
N.check(N.lib().pl_orb_set_profiling(ex._h, 1))
for _ in range(a.iters):
    run()
ex.sync()
out = np.zeros(5, np.float32)
ch = C.c_int()
N.check(N.lib().pl_orb_stage_ms(ex._h, N.ptr(out), C.byref(ch)))
P, Pb = C.c_longlong(), C.c_longlong()
N.check(N.lib().pl_orb_bytes_per_frame(ex._h, None, None, C.byref(P), C.byref(Pb)))
P, Pb = P.value, Pb.value
nframes = a.frames * a.iters
K = float(d_n.float().mean().item())
alg = [a.w * a.h + (P - 0) + Pb, Pb + 16 * 0, 0, 2 * P, 749 * K + 4 * K + 512 * K + 32 * K]
names = ["pyramid", "fast_cells", "octree", "blur7", "orient_brief"]
for i in range(5):
    us = out[i] * 1000 / nframes
    gbs = alg[i] / (us * 1e-6) / 1e9 if alg[i] and us > 0 else 0
    print(f"  {names[i]:13s} {us:8.2f} us/frame   alg {alg[i]/1e6:6.2f} MB/frame  {gbs:8.1f} GB/s")
print("mean kps/frame", K, "launches", ex.last_launches())
