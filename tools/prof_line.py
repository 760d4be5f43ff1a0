"""Stage timing of the batched line path (LSD + LBD) with inputs resident in HBM."""
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
ap.add_argument("--frames", type=int, default=64)
ap.add_argument("--chunk", type=int, default=64)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--w", type=int, default=640)
ap.add_argument("--h", type=int, default=480)
ap.add_argument("--room", action="store_true", help="frames of the bench's room sequence instead of the textured test frames")
a = ap.parse_args()
fr = pkg.synth.room_sequence(a.frames, a.w, a.h, workers=min(32, os.cpu_count() or 1))[0] if a.room else pkg.synth.frames(6000, a.frames, a.w, a.h)
d_in = torch.from_numpy(fr).cuda()
ex = api.LineExtractor(max_cols=a.w, max_rows=a.h, max_batch=a.chunk)
ML = 80
d_kls = torch.empty((a.frames, ML, 17), dtype=torch.float32, device="cuda")
d_desc = torch.empty((a.frames, ML, 32), dtype=torch.uint8, device="cuda")
d_co = torch.empty((a.frames, ML, 3), dtype=torch.float64, device="cuda")
d_n = torch.empty(a.frames, dtype=torch.int32, device="cuda")
st = torch.cuda.ExternalStream(ex.stream())
torch.cuda.synchronize()


def run():
    ex.extract_batch_dev(d_in.data_ptr(), a.frames, a.h, a.w, a.w, a.w * a.h, ML, d_kls.data_ptr(), d_desc.data_ptr(), d_co.data_ptr(), d_n.data_ptr())


for _ in range(2):
    run()
ex.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(a.iters):
    run()
e1.record(st)
ex.sync()
ms = e0.elapsed_time(e1) / a.iters
print(f"total: {ms:.3f} ms / {a.frames} frames = {ms*1000/a.frames:.1f} us/frame = {a.frames/ms*1000:.0f} frames/s (chunk {a.chunk})")
N.check(N.lib().pl_line_set_profiling(ex._h, 1))
for _ in range(a.iters):
    run()
ex.sync()
out = np.zeros(5, np.float32)
ch = C.c_int()
N.check(N.lib().pl_line_stage_ms(ex._h, N.ptr(out), C.byref(ch)))
names = ["scale+grad", "seed_sort", "grow+nfa", "keylines+sobel", "lbd"]
for i in range(5):
    print(f"  {names[i]:15s} {out[i]*1000/(a.frames*a.iters):9.2f} us/frame  ({out[i]/a.iters:.3f} ms per pass)")
nfa = C.c_float()
N.check(N.lib().pl_line_nfa_ms(ex._h, C.byref(nfa)))
print(f"  of grow+nfa, k_lsd_nfa after the grower: {nfa.value / a.iters:.3f} ms per pass")
print("lines/frame", float(d_n.float().mean().item()), "launches", ex.last_launches())
ph = np.zeros(16, np.int64)
N.check(N.lib().pl_line_grow_phases(ex._h, C.c_int(0), N.ptr(ph)))
M = 1e6
print(f"grow kernel (frame 0): frame active {ph[1]/M:.1f} Mcycles; sequencer: commit {ph[2]/M:.1f}, re-growth {ph[0]/M:.1f}, issue {ph[8]/M:.1f}, "
      f"idle {ph[9]/M:.1f}; {ph[13]} growers: growing {ph[10]/M:.1f} (given up {ph[14]/M:.1f}), waiting {ph[11]/M:.1f}, parking {ph[12]/M:.1f} "
      f"(summed over the warps); tickets {ph[3]}, committed {ph[5]}, void {ph[7]}, deferred {ph[6]}, regrown {ph[4]}")
