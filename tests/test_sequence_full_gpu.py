"""GPU parity at the size the bench runs (-m gpu): the whole 300-frame sequence of BASELINE.json config 2 — every key point, ORB
descriptor, key line and LBD descriptor of every frame, and the summary of all four searches (C3, C2 through
Tracking::SearchLocalPoints in one call, D3, D5) frame by frame — CUDA path through the C ABI against the CPU oracle.
The oracle extracts frame-parallel on the host threads (ctypes releases the GIL); its matching is sequential."""
import concurrent.futures
import importlib
import os
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PKG = "orb_slam2_modification_with-point-and-line-feature_b200"
N_FRAMES = int(os.environ.get("PLSLAM_FULL_SEQUENCE_FRAMES", "300"))


def test_bench_sequence_equals_the_oracle_frame_by_frame(synth, api, oracle):
    fe = importlib.import_module(PKG + ".frontend")
    threads = max(2, min(32, os.cpu_count() or 2))
    gray, depth, T = synth.room_sequence(N_FRAMES, 640, 480, workers=threads)
    ob = oracle.OracleBackend(1000)
    sf = ob.scale_factors()
    local = threading.local()

    def orb_job(i):
        if not hasattr(local, "o"):
            local.o = oracle.OrbOracle(1000)
        return local.o.extract(gray[i])

    with concurrent.futures.ThreadPoolExecutor(threads) as ex:
        fo = [ex.submit(orb_job, i) for i in range(N_FRAMES)]
        fl = [ex.submit(oracle.line_extract, gray[i], 80) for i in range(N_FRAMES)]
        feats_o = ([f.result() for f in fo], [f.result() for f in fl])
    gb = fe.GpuBackend(api, 480, 640, 1000, chunk=N_FRAMES)
    feats_g = (gb.extract_orb(gray), gb.extract_lines(gray))
    # ---- extraction: every frame ----
    n_kp = n_kl = 0
    for t in range(N_FRAMES):
        (ko, do), (kg, dg) = feats_o[0][t], feats_g[0][t]
        assert len(kg) == len(ko) and np.array_equal(dg, do), f"ORB, frame {t}"
        for f in ("x", "y", "angle", "response", "octave", "size"):
            assert np.array_equal(kg[f], ko[f]), (t, f)
        (lo, ldo, lco), (lg, ldg, lcg) = feats_o[1][t], feats_g[1][t]
        assert len(lg) == len(lo) and np.array_equal(lg["class_id"], lo["class_id"]) and np.array_equal(ldg, ldo), f"lines, frame {t}"
        for f in ("sx", "sy", "ex", "ey"):
            assert np.abs(lg[f] - lo[f]).max(initial=0) <= 1e-3, (t, f)
        assert np.allclose(lcg, lco, rtol=0, atol=1e-9)
        n_kp += len(kg)
        n_kl += len(lg)
    assert n_kp > 900 * N_FRAMES and n_kl > 60 * N_FRAMES
    # ---- the matching schedule of the bench, each arm on ITS OWN features ----
    so = fe.TrackingFrontEnd(ob, device_glue=True).run(gray, depth, T, sf, features=feats_o)
    sg = fe.TrackingFrontEnd(gb, device_glue=True).run(gray, depth, T, sf, features=feats_g)
    assert len(sg) == len(so) == N_FRAMES
    for t, (a, b) in enumerate(zip(sg, so)):
        assert a == b, f"frame {t}: {a} != {b}"
    late = so[N_FRAMES // 2:]
    assert sum(r.get("c2_matches", 0) for r in late) > 50 * len(late) and sum(r.get("d5_matches", 0) for r in late) > 20 * len(late)
