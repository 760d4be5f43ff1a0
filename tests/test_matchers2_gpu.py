"""GPU parity tests (-m gpu) for the remaining matcher rows of SURVEY.md §8 through the C ABI vs the oracle:
C4 (relocalisation projection), C5 (Sim3 projection), C6 / C7 (SearchByBoW), D6 (brute-force line matchers).
Bar: every index and count bit-exact."""
import importlib

import numpy as np
import pytest

import matchgen

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,n,m,th,od", [(1, 1000, 900, 10.0, 100), (2, 1000, 2000, 3.0, 64), (3, 40, 600, 15.0, 100), (4, 0, 30, 10.0, 100),
                                             (5, 500, 0, 10.0, 100)])
def test_keyframe_points_c4(seed, n, m, th, od, api, oracle, synth):
    N = api.N
    sf = oracle.OrbOracle().tables()["scale_factors"]
    rng = np.random.default_rng(seed)
    fv, pv, ow, log_sf, keep = matchgen.pose_case(rng, n, m, N, synth.TUM1, sf, 2)
    dm = api.DescriptorMatcher()
    for ori in (True, False):
        g = dm.SearchByProjectionKeyFrameBatch([fv], [pv], ow[None], [log_sf], th, od, ori)[0]
        o = oracle.search_keyframe_points(fv, pv, ow, log_sf, th, od, ori)
        assert np.array_equal(g[0], o[0]) and g[1] == o[1]
    if n >= 1000 and m >= 900:
        assert o[1] > 100


@pytest.mark.parametrize("seed,n,m,th", [(11, 1000, 900, 10), (12, 1500, 3000, 4), (13, 30, 500, 10), (14, 0, 10, 10), (15, 200, 0, 10)])
def test_sim3_points_c5(seed, n, m, th, api, oracle, synth):
    N = api.N
    sf = oracle.OrbOracle().tables()["scale_factors"]
    rng = np.random.default_rng(seed)
    fv, pv, ow, log_sf, keep = matchgen.pose_case(rng, n, m, N, synth.TUM1, sf, 3, claimed_frac=0.3)
    g = api.DescriptorMatcher().SearchByProjectionSim3Batch([fv], [pv], ow[None], [log_sf], th)[0]
    o = oracle.search_sim3_points(fv, pv, ow, log_sf, th)
    assert np.array_equal(g[0], o[0]) and g[1] == o[1]
    if n >= 1000:
        assert o[1] > 50


def test_pose_searches_batched_ragged(api, oracle, synth):
    N = api.N
    sf = oracle.OrbOracle().tables()["scale_factors"]
    rng = np.random.default_rng(99)
    cases = [matchgen.pose_case(rng, n, m, N, synth.TUM1, sf, 2) for n, m in [(800, 700), (0, 5), (300, 0), (1200, 1500)]]
    dm = api.DescriptorMatcher()
    g = dm.SearchByProjectionKeyFrameBatch([c[0] for c in cases], [c[1] for c in cases], np.stack([c[2] for c in cases]), [c[3] for c in cases], 8.0,
                                           90, True)
    for c, gi in zip(cases, g):
        o = oracle.search_keyframe_points(c[0], c[1], c[2], c[3], 8.0, 90, True)
        assert np.array_equal(gi[0], o[0]) and gi[1] == o[1]
    g5 = dm.SearchByProjectionSim3Batch([c[0] for c in cases], [c[1] for c in cases], np.stack([c[2] for c in cases]), [c[3] for c in cases], 10)
    for c, gi in zip(cases, g5):
        o = oracle.search_sim3_points(c[0], c[1], c[2], c[3], 10)
        assert np.array_equal(gi[0], o[0]) and gi[1] == o[1]


@pytest.mark.parametrize("seed,nA,nB,nodes,mode", [(1, 1000, 1000, 100, 0), (2, 1000, 1000, 100, 1), (3, 2000, 1500, 12, 0), (4, 64, 900, 3, 1),
                                                   (5, 0, 50, 4, 0), (6, 50, 0, 4, 1), (7, 700, 700, 700, 0)])
def test_search_by_bow_c6_c7(seed, nA, nB, nodes, mode, api, oracle):
    N = api.N
    rng = np.random.default_rng(seed)
    a, b, keep = matchgen.bow_case(rng, nA, nB, nodes, N, mode)
    dm = api.DescriptorMatcher()
    for ori in (True, False):
        for ratio in (0.75, 0.9):
            g = dm.SearchByBoWBatch([a], [b], mode, ratio, ori)[0]
            o = oracle.search_bow(a, b, mode, ratio, ori)
            assert np.array_equal(g[0], o[0]) and g[1] == o[1]
    if nA >= 1000 and nB >= 1000:
        assert o[1] > 150


def test_search_by_bow_feature_in_two_nodes(api, oracle):
    """The nodes of a call are replayed side by side because a feature belongs to one node.  A caller-built view that lists a
    feature under two nodes makes them depend on each other: the library detects it and replays the call in one ordered pass."""
    N = api.N
    rng = np.random.default_rng(11)
    n = 600
    desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    descb = matchgen.noisy(desc, rng, 0.05)
    ang = rng.uniform(0, 360, n).astype(np.float32)
    nodes = rng.integers(0, 20, n)
    fva = {int(k): [int(i) for i in np.nonzero(nodes == k)[0]] for k in np.unique(nodes)}
    fvb = {k: list(v) for k, v in fva.items()}
    for k in list(fvb)[:-1]:  # every node of side B also lists the first features of the next node
        nxt = sorted(fvb)[sorted(fvb).index(k) + 1]
        fvb[k] = sorted(set(fvb[k]) | set(fva[nxt][:5]))
    keep = []
    a = N.make_bow_view(ang, desc, None, fva, keep)
    b = N.make_bow_view(ang, descb, None, fvb, keep)
    for mode in (0, 1):
        g = api.DescriptorMatcher().SearchByBoWBatch([a], [b], mode, 0.9, True)[0]
        o = oracle.search_bow(a, b, mode, 0.9, True)
        assert np.array_equal(g[0], o[0]) and g[1] == o[1]


def test_search_by_bow_batched(api, oracle):
    N = api.N
    rng = np.random.default_rng(5)
    cases = [matchgen.bow_case(rng, nA, nB, nodes, N, 0) for nA, nB, nodes in [(500, 600, 50), (0, 10, 3), (900, 900, 90), (10, 0, 3)]]
    g = api.DescriptorMatcher().SearchByBoWBatch([c[0] for c in cases], [c[1] for c in cases], 0, 0.7, True)
    for c, gi in zip(cases, g):
        o = oracle.search_bow(c[0], c[1], 0, 0.7, True)
        assert np.array_equal(gi[0], o[0]) and gi[1] == o[1]


@pytest.mark.parametrize("seed,nr,nc", [(1, 80, 80), (2, 300, 2000), (3, 5, 1), (4, 0, 10), (5, 1500, 1500)])
def test_line_bruteforce_d6(seed, nr, nc, api, oracle):
    rng = np.random.default_rng(seed)
    cur = rng.integers(0, 256, (nc, 32), dtype=np.uint8)
    ref = matchgen.noisy(cur[rng.integers(0, nc, nr)], rng, 0.06) if nc and nr else rng.integers(0, 256, (nr, 32), dtype=np.uint8)
    if nr > 10:
        ref[:: 9] = rng.integers(0, 256, ref[:: 9].shape, dtype=np.uint8)
        ref[1] = cur[0]                                                   # an exact duplicate: distance 0
    dm = api.DescriptorMatcher()
    g, o = dm.LineMatchKnnRatio(ref, cur), oracle.line_match_knn_ratio(ref, cur)
    assert np.array_equal(g[0], o[0]) and g[1] == o[1]
    g, o = dm.LineSearchForTriangulation(ref, cur), oracle.line_search_for_triangulation(ref, cur)
    assert np.array_equal(g[0], o[0]) and g[1] == o[1] and g[2] == o[2]
    valid = (rng.random(nr) < 0.8).astype(np.uint8)
    for va in (valid, None):
        g, o = dm.LineFuseCandidates(ref, va, cur), oracle.line_fuse_candidates(ref, va, cur)
        assert np.array_equal(g[0], o[0]) and g[1] == o[1]
