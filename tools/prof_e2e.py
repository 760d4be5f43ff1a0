"""Timeline of the end-to-end step of bench.py (host images -> features -> matches, caller glue included).  Run on a GPU box."""
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "orb_slam2_modification_with-point-and-line-feature_b200"
pkg = importlib.import_module(PKG)
api = pkg.load_api()
fe = importlib.import_module(PKG + ".frontend")
N = api.N
F = int(sys.argv[1]) if len(sys.argv) > 1 else 300
H, W, MAXL = 480, 640, 80
gray, depth, Tcw = pkg.synth.room_sequence(F, W, H, workers=min(32, os.cpu_count() or 1))
gb = fe.GpuBackend(api, H, W, 1000, chunk=F, device=0)
RES = int(os.environ.get("RESERVE_SMS", "16"))
DG = int(os.environ.get("DEVICE_GLUE", "1"))
gb.line.set_reserved_sms(RES)
sf = gb.scale_factors()
cap = gb.orb.max_keypoints()
h_gray = torch.from_numpy(gray).pin_memory()
e_gray = torch.empty(gray.shape, dtype=torch.uint8, device="cuda")
d_kps = torch.empty((F, cap, 7), dtype=torch.float32, device="cuda")
d_desc = torch.empty((F, cap, 32), dtype=torch.uint8, device="cuda")
d_n = torch.empty(F, dtype=torch.int32, device="cuda")
d_kls = torch.empty((F, MAXL, 17), dtype=torch.float32, device="cuda")
d_ldesc = torch.empty((F, MAXL, 32), dtype=torch.uint8, device="cuda")
d_lco = torch.empty((F, MAXL, 3), dtype=torch.float64, device="cuda")
d_ln = torch.empty(F, dtype=torch.int32, device="cuda")
p_kps, p_desc, p_n = (torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in (d_kps, d_desc, d_n))
p_kls, p_ldesc, p_lco, p_ln = (torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in (d_kls, d_ldesc, d_lco, d_ln))
s_orb, s_line = torch.cuda.ExternalStream(gb.orb.stream()), torch.cuda.ExternalStream(gb.line.stream())
ev_orb = torch.cuda.Event()
marks = []


def mark(name):
    marks.append((name, time.perf_counter()))


class Timed:
    def __init__(self, b):
        self.b = b

    def __getattr__(self, k):
        f = getattr(self.b, k)
        if k.endswith("_batch"):
            def w(*a):
                mark(k + " >")
                r = f(*a)
                mark(k + " <")
                return r
            return w
        return f


class LinesLater:
    def result(self):
        mark("lines wait >")
        gb.line.sync()
        mark("lines ready")
        for dst, src in ((p_kls, d_kls), (p_ldesc, d_ldesc), (p_lco, d_lco), (p_ln, d_ln)):
            dst.copy_(src, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        kl = p_kls.numpy().view(N.KL_DTYPE).reshape(F, MAXL)
        ld, lc, ln = p_ldesc.numpy(), p_lco.numpy(), p_ln.numpy()
        r = [(kl[i, :ln[i]], ld[i, :ln[i]], lc[i, :ln[i]]) for i in range(F)]
        mark("lines on host")
        return r


tfe = fe.TrackingFrontEnd(Timed(gb), device_glue=bool(DG))


def step():
    marks.clear()
    mark("start")
    e_gray.copy_(h_gray, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    mark("images on device")
    gb.orb.extract_batch_dev(e_gray.data_ptr(), F, H, W, W, W * H, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_n.data_ptr())
    ev_orb.record(s_orb)
    s_line.wait_event(ev_orb)
    gb.line.extract_batch_dev(e_gray.data_ptr(), F, H, W, W, W * H, MAXL, d_kls.data_ptr(), d_ldesc.data_ptr(), d_lco.data_ptr(), d_ln.data_ptr())
    gb.orb.sync()
    mark("ORB done")
    for dst, src in ((p_kps, d_kps), (p_desc, d_desc), (p_n, d_n)):
        dst.copy_(src, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    kp = p_kps.numpy().view(N.KP_DTYPE).reshape(F, cap)
    dd, nn = p_desc.numpy(), p_n.numpy()
    orb = [(kp[i, :nn[i]], dd[i, :nn[i]]) for i in range(F)]
    mark("ORB on host")
    s = tfe.run(gray, depth, Tcw, sf, features=(orb, LinesLater()))
    mark("end")
    return s


step()
step()
for _ in range(2):
    step()
    t0 = marks[0][1]
    print(" | ".join("%s %.1f" % (n, 1e3 * (t - t0)) for n, t in marks))
