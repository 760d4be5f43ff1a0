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


class DenseResultOracle(ThreadedOracle):
    """line_search_batch answers the way the CUDA backend does: a sequence that carries the dense (n, max lines) match array."""

    def line_search_batch(self, cvs, lvs, second=False):
        api = importlib.import_module(PKG + ".api")
        res = self.ob.line_search_batch(cvs, lvs)
        nl = np.array([len(r[0]) for r in res], np.int64)
        out = np.full((len(res), max(int(nl.max()) if len(nl) else 1, 1)), -1, np.int32)
        for i, r in enumerate(res):
            out[i, :nl[i]] = r[0]
        return api._LineMatches(out, nl, *(np.array([r[k] for r in res], np.int32) for k in (1, 2, 3)))


def test_dense_paths_with_frames_and_key_frames_without_lines(seq, feats, oracle):
    """Frames without a single line (one of them a key frame) and a key frame none of whose lines has depth at both ends: the
    snapshots of the local line map as slices of one gathered array, the D5 views built from them and the dense result summaries
    must give what the per-frame path gives on the same features."""
    fe = importlib.import_module(PKG + ".frontend")
    N = importlib.import_module(PKG + "._native")
    gray, depth, T = seq
    depth = depth.copy()
    depth[8] = 0.0           # key frame 8 (a key frame every 4 frames here) lifts nothing: its lines (and points) have no depth
    ob = oracle.OracleBackend(1000)
    sf = ob.scale_factors()
    orb_l, line_l = feats
    line_l = list(line_l)
    for t in (0, 3):          # key frame 0 and frame 3 have no lines at all
        line_l[t] = (line_l[t][0][:0], line_l[t][1][:0], line_l[t][2][:0])
    ref = fe.TrackingFrontEnd(ob, keyframe_every=4).run(gray, depth, T, sf, features=(orb_l, line_l), batch=True)
    assert sum(r.get("d5_matches", 0) for r in ref) > 0 and "d3_matches" not in ref[3] and "d3_matches" not in ref[4]
    n = len(gray)
    ml = 80
    kl = np.zeros((n, ml), N.KL_DTYPE)
    ld = np.zeros((n, ml, 32), np.uint8)
    lc = np.zeros((n, ml, 3), np.float64)
    ln = np.array([len(l[0]) for l in line_l], np.int32)
    for i in range(n):
        kl[i, :ln[i]], ld[i, :ln[i]], lc[i, :ln[i]] = line_l[i][0], line_l[i][1], line_l[i][2]
    lines = fe.FeatureList((kl[i, :ln[i]], ld[i, :ln[i]], lc[i, :ln[i]]) for i in range(n))
    lines.dense = (kl, ln, ld)
    for backend in (ThreadedOracle(ob), DenseResultOracle(ob)):
        assert fe.TrackingFrontEnd(backend, keyframe_every=4).run(gray, depth, T, sf, features=(orb_l, lines), batch=True) == ref


def test_line_matches_sequence_behaves_like_the_list_it_replaces():
    api = importlib.import_module(PKG + ".api")
    out = np.array([[4, -1, 7], [2, -1, -1]], np.int32)
    r = api._LineMatches(out, np.array([3, 1]), np.array([2, 1], np.int32), np.array([0, 1], np.int32), np.array([9, 5], np.int32))
    assert len(r) == 2 and r[1][1:] == (1, 1, 5) and r[-1][0].tolist() == [2] and [x[0].tolist() for x in r] == [[4, -1, 7], [2]]
    assert [x[1] for x in r[0:2]] == [2, 1]
    with pytest.raises(IndexError):
        r[2]
