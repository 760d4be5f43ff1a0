"""The other configurations of BASELINE.json, measured on one GPU — they are parity-test cases, not bench lines,
so this writes a small JSON of its own (gpurun_out/configs.json; the committed copy lives in profiles/).

  config 1  one 640x480 frame: ORB(1000) + LSD/LBD(80), GPU latency through the host-pointer C ABI vs the CPU oracle (1 thread)
  config 3  KITTI-size 1241x376 frames, ORB(2000) + lines, per-frame latency p50 / p95 (both extractors on their own streams)
  config 4  brute-force Hamming kNN-2 sweep, N = M = 1k .. 16k, device-resident; checked against the oracle up to 4k
  config 5  batched offline extraction of 4096 frames (ORB + lines), frames/s on this rank + a checksum of all outputs
"""
import hashlib
import importlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
pkg = importlib.import_module("orb_slam2_modification_with-point-and-line-feature_b200")
api = pkg.load_api()
import pyoracle  # the checker / CPU leg only

N5 = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
out = {}


def pct(v, p):
    return float(np.percentile(np.asarray(v), p))


# ---------------- config 1 ----------------
img = pkg.synth.frame(1000, 640, 480)
ex, lx = api.ORBextractor(1000, 1.2, 8, 20, 7), api.LineExtractor()
for _ in range(3):
    ex(img), lx.ExtractLineSegment(img)
t_orb, t_line = [], []
for _ in range(20):
    t0 = time.perf_counter(); k, d = ex(img); t1 = time.perf_counter(); kl, ld, lc = lx.ExtractLineSegment(img); t2 = time.perf_counter()
    t_orb.append(1e3 * (t1 - t0)); t_line.append(1e3 * (t2 - t1))
oo = pyoracle.OrbOracle(1000)
t0 = time.perf_counter(); ok, od = oo.extract(img); t1 = time.perf_counter(); okl, old_, olc = pyoracle.line_extract(img, 80); t2 = time.perf_counter()
assert np.array_equal(d, od) and len(kl) == len(okl) and np.array_equal(ld, old_)
out["config1_single_frame_640x480"] = {"gpu_orb_ms_p50": pct(t_orb, 50), "gpu_line_ms_p50": pct(t_line, 50), "cpu_oracle_orb_ms": 1e3 * (t1 - t0),
                                       "cpu_oracle_line_ms": 1e3 * (t2 - t1), "keypoints": int(len(k)), "lines": int(len(kl)),
                                       "parity": "ORB descriptors and LBD bits equal to the oracle"}
ex.close(); lx.close()

# ---------------- config 3 ----------------
W3, H3 = 1241, 376
fr3 = pkg.synth.frames(3000, 24, W3, H3)
ex, lx = api.ORBextractor(2000, 1.2, 8, 20, 7, max_cols=W3, max_rows=H3), api.LineExtractor(max_cols=W3, max_rows=H3)
d_in = torch.from_numpy(fr3).cuda()
cap = ex.max_keypoints()
d_kps = torch.empty((1, cap, 7), dtype=torch.float32, device="cuda"); d_desc = torch.empty((1, cap, 32), dtype=torch.uint8, device="cuda")
d_n = torch.empty(1, dtype=torch.int32, device="cuda")
d_kls = torch.empty((1, 80, 17), dtype=torch.float32, device="cuda"); d_ld = torch.empty((1, 80, 32), dtype=torch.uint8, device="cuda")
d_co = torch.empty((1, 80, 3), dtype=torch.float64, device="cuda"); d_ln = torch.empty(1, dtype=torch.int32, device="cuda")
lat = []
for i in range(24):
    p = d_in[i].data_ptr()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ex.extract_batch_dev(p, 1, H3, W3, W3, W3 * H3, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_n.data_ptr())
    lx.extract_batch_dev(p, 1, H3, W3, W3, W3 * H3, 80, d_kls.data_ptr(), d_ld.data_ptr(), d_co.data_ptr(), d_ln.data_ptr())
    ex.sync(); lx.sync()
    if i >= 4:
        lat.append(1e3 * (time.perf_counter() - t0))
out["config3_kitti_1241x376_orb2000_lines"] = {"frames": len(lat), "latency_ms_p50": pct(lat, 50), "latency_ms_p95": pct(lat, 95),
                                               "note": "one frame at a time, ORB || LSD+LBD on their own streams, image resident in HBM"}
ex.close(); lx.close()

# ---------------- config 4 ----------------
dm = api.DescriptorMatcher()
sweep = []
for n in (1024, 2048, 4096, 8192, 16384):
    q, t = pkg.synth.descriptor_sets(n)
    dq, dt = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    di = torch.empty((len(q), 2), dtype=torch.int32, device="cuda"); dd = torch.empty((len(q), 2), dtype=torch.int32, device="cuda")
    for _ in range(3):
        dm.knn2_dev(dq.data_ptr(), len(q), dt.data_ptr(), len(t), di.data_ptr(), dd.data_ptr())
    dm.sync()
    st = torch.cuda.ExternalStream(dm.stream())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(10):
        dm.knn2_dev(dq.data_ptr(), len(q), dt.data_ptr(), len(t), di.data_ptr(), dd.data_ptr())
    e1.record(st)
    dm.sync()
    ms = e0.elapsed_time(e1) / 10
    row = {"n": n, "ms": ms, "pairs_per_s": len(q) * len(t) / (ms * 1e-3), "popc32_per_s": 8.0 * len(q) * len(t) / (ms * 1e-3)}
    if n <= 4096:
        oi, od2 = pyoracle.hamming_knn2(q, t)
        row["bit_exact_vs_oracle"] = bool(np.array_equal(di.cpu().numpy(), oi) and np.array_equal(dd.cpu().numpy(), od2))
        assert row["bit_exact_vs_oracle"]
    sweep.append(row)
out["config4_hamming_knn2_sweep"] = sweep

# ---------------- config 5 ----------------
CH = 512
ex, lx = api.ORBextractor(1000, 1.2, 8, 20, 7, max_batch=CH), api.LineExtractor(max_batch=CH)
cap = ex.max_keypoints()
base = pkg.synth.frames(6000, CH)          # frames 6000 .. 6000+CH-1; the other chunks are these frames shifted by a few columns
d_chunk = torch.from_numpy(base).cuda()
d_kps = torch.empty((CH, cap, 7), dtype=torch.float32, device="cuda"); d_desc = torch.empty((CH, cap, 32), dtype=torch.uint8, device="cuda")
d_n = torch.empty(CH, dtype=torch.int32, device="cuda")
d_kls = torch.empty((CH, 80, 17), dtype=torch.float32, device="cuda"); d_ld = torch.empty((CH, 80, 32), dtype=torch.uint8, device="cuda")
d_co = torch.empty((CH, 80, 3), dtype=torch.float64, device="cuda"); d_ln = torch.empty(CH, dtype=torch.int32, device="cuda")
s_orb, s_line = torch.cuda.ExternalStream(ex.stream()), torch.cuda.ExternalStream(lx.stream())
ev = torch.cuda.Event()
h = hashlib.sha256()
n_chunks = max(1, N5 // CH)
tot_kp = tot_kl = 0
t_total = 0.0
for c in range(n_chunks + 1):            # chunk 0 is the warm-up
    cur = torch.roll(d_chunk, shifts=3 * c, dims=2).contiguous()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ex.extract_batch_dev(cur.data_ptr(), CH, 480, 640, 640, 640 * 480, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_n.data_ptr())
    ev.record(s_orb)
    s_line.wait_event(ev)
    lx.extract_batch_dev(cur.data_ptr(), CH, 480, 640, 640, 640 * 480, 80, d_kls.data_ptr(), d_ld.data_ptr(), d_co.data_ptr(), d_ln.data_ptr())
    ex.sync(); lx.sync()
    dt_ = time.perf_counter() - t0
    if c == 0:
        continue
    t_total += dt_
    nn, ln = d_n.cpu().numpy(), d_ln.cpu().numpy()
    tot_kp += int(nn.sum()); tot_kl += int(ln.sum())
    h.update(nn.tobytes()); h.update(ln.tobytes())
    h.update(d_desc.cpu().numpy()[np.arange(cap)[None, :] < nn[:, None]].tobytes())
    h.update(d_ld.cpu().numpy()[np.arange(80)[None, :] < ln[:, None]].tobytes())
out["config5_batched_extraction"] = {"frames": n_chunks * CH, "chunk": CH, "frames_per_s": n_chunks * CH / t_total, "seconds": t_total,
                                     "keypoints": tot_kp, "lines": tot_kl, "sha256_counts_descriptors": h.hexdigest(),
                                     "note": "device-resident chunks of 512 frames (the 512 textured test frames, shifted per chunk), ORB then LSD+LBD; "
                                             "hash over counts, ORB descriptors and LBD descriptors of every frame"}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
