"""CPU tests (-m "not gpu") of the caller glue (frontend.TrackingFrontEnd) that drives both arms of the bench: the schedule must
give the same summary whether the searches are issued frame by frame, batched, or batched with the point side and the two line
searches on their own host threads (what GpuBackend does), and whether the Frame glue runs in numpy or through the batched F-row
calls.  The oracle backend stands in for the device here: the host logic is what is under test."""
import importlib

import numpy as np
import pytest

PKG = "orb_slam2_modification_with-point-and-line-feature_b200"


@pytest.fixture(scope="module")
def seq(synth):
    return synth.room_sequence(14, 640, 480, workers=4)


@pytest.fixture(scope="module")
def feats(seq, oracle):
    ob = oracle.OracleBackend(1000)
    return ob.extract_orb(seq[0]), ob.extract_lines(seq[0])


class ThreadedOracle:
    """The oracle backend behind GpuBackend's interface for concurrent sides (second line-matcher 'handle' included)."""

    concurrent_sides = True

    def __init__(self, ob):
        self.ob = ob
        self.second_calls = 0

    def __getattr__(self, k):
        return getattr(self.ob, k)

    def line_search_batch(self, cvs, lvs, second=False):
        self.second_calls += int(second)
        return self.ob.line_search_batch(cvs, lvs)


@pytest.mark.parametrize("device_glue", [False, True])
def test_schedule_is_independent_of_batching_and_threads(device_glue, seq, feats, oracle):
    fe = importlib.import_module(PKG + ".frontend")
    gray, depth, T = seq
    ob = oracle.OracleBackend(1000)
    sf = ob.scale_factors()
    ref = fe.TrackingFrontEnd(ob, device_glue=device_glue).run(gray, depth, T, sf, features=feats, batch=False)
    assert fe.TrackingFrontEnd(ob, device_glue=device_glue).run(gray, depth, T, sf, features=feats, batch=True) == ref
    tb = ThreadedOracle(ob)
    for _ in range(3):   # the threaded schedule is deterministic
        assert fe.TrackingFrontEnd(tb, device_glue=device_glue).run(gray, depth, T, sf, features=feats, batch=True) == ref
    assert tb.second_calls == 3      # D5 went to the second handle each time
    assert sum(r.get("c3_matches", 0) for r in ref) > 3000 and sum(r.get("d3_matches", 0) for r in ref) > 100


def test_worker_errors_surface_on_the_callers_thread(seq, feats, oracle):
    fe = importlib.import_module(PKG + ".frontend")
    gray, depth, T = seq
    ob = oracle.OracleBackend(1000)

    class Failing(ThreadedOracle):
        def search_last_frame_batch(self, *a):
            raise RuntimeError("boom")

    with pytest.raises(RuntimeError, match="boom"):
        fe.TrackingFrontEnd(Failing(ob)).run(gray, depth, T, ob.scale_factors(), features=feats, batch=True)


def test_dense_feature_arrays_give_the_same_schedule(seq, feats, oracle):
    """The extractors' own (n, capacity) output arrays (what the end-to-end leg of bench.py hands over) take the vectorised paths of
    the glue — one pass over the sequence for the line endpoints, the views of the line searches built as structured arrays: the
    summary must be the one of the per-frame lists."""
    fe = importlib.import_module(PKG + ".frontend")
    N = importlib.import_module(PKG + "._native")
    gray, depth, T = seq
    ob = oracle.OracleBackend(1000)
    sf = ob.scale_factors()
    ref = fe.TrackingFrontEnd(ob).run(gray, depth, T, sf, features=feats, batch=True)
    orb_l, line_l = feats
    n = len(gray)
    cap = max(len(o[0]) for o in orb_l) + 7
    kp = np.zeros((n, cap), N.KP_DTYPE)
    dd = np.full((n, cap, 32), 0xAB, np.uint8)
    nn = np.array([len(o[0]) for o in orb_l], np.int32)
    ml = max(len(l[0]) for l in line_l) + 3
    kl = np.zeros((n, ml), N.KL_DTYPE)
    for k in ("sx", "sy", "ex", "ey"):
        kl[k] = np.nan     # rows beyond a frame's count hold anything
    ld = np.full((n, ml, 32), 0xCD, np.uint8)
    lc = np.zeros((n, ml, 3), np.float64)
    ln = np.array([len(l[0]) for l in line_l], np.int32)
    for i in range(n):
        kp[i, :nn[i]], dd[i, :nn[i]] = orb_l[i][0], orb_l[i][1]
        kl[i, :ln[i]], ld[i, :ln[i]], lc[i, :ln[i]] = line_l[i][0], line_l[i][1], line_l[i][2]
    orb = fe.FeatureList((kp[i, :nn[i]], dd[i, :nn[i]]) for i in range(n))
    orb.dense = (kp, nn)
    lines = fe.FeatureList((kl[i, :ln[i]], ld[i, :ln[i]], lc[i, :ln[i]]) for i in range(n))
    lines.dense = (kl, ln, ld)
    tfe = fe.TrackingFrontEnd(ThreadedOracle(ob))
    assert tfe.run(gray, depth, T, sf, features=(orb, lines), batch=True) == ref
    assert tfe.keepalive and any(isinstance(k, dict) and "s3" in k for k in tfe.keepalive)   # the dense path was the one taken
