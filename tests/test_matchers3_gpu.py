"""GPU parity tests (-m gpu) for the E rows through the C ABI vs the oracle: ORBmatcher::Fuse (both
overloads), SearchBySim3, SearchForInitialization, SearchForTriangulation and MapPoint / MapLine::ComputeDistinctiveDescriptors.
Bar: every index, distance and count bit-exact."""
import numpy as np
import pytest

import matchgen

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,m,extra,th", [(1, 900, 300, 3.0), (2, 2500, 500, 4.0), (3, 30, 0, 3.0), (4, 0, 50, 3.0), (5, 400, 2000, 2.5)])
def test_fuse_candidates(seed, m, extra, th, api, oracle, synth):
    N = api.N
    sf = oracle.OrbOracle().tables()["scale_factors"]
    rng = np.random.default_rng(seed)
    fv, pv, ow, log_sf, inv_s2, keep = matchgen.fuse_case(rng, m, extra, N, synth.TUM1, sf)
    dm = api.DescriptorMatcher()
    for variant in (0, 1):
        g = dm.FuseCandidatesBatch([fv], [pv], ow[None], [log_sf], [inv_s2] if variant == 0 else None, th, variant)[0]
        o = oracle.fuse_candidates(fv, pv, ow, log_sf, inv_s2, th, variant)
        assert np.array_equal(g[0], o[0]) and np.array_equal(g[1], o[1]) and g[2] == o[2]
    if m >= 900:
        assert o[2] > 200


def test_fuse_candidates_batched_ragged(api, oracle, synth):
    N = api.N
    sf = oracle.OrbOracle().tables()["scale_factors"]
    rng = np.random.default_rng(77)
    cases = [matchgen.fuse_case(rng, m, e, N, synth.TUM1, sf) for m, e in [(700, 100), (0, 5), (5, 0), (1500, 300)]]
    g = api.DescriptorMatcher().FuseCandidatesBatch([c[0] for c in cases], [c[1] for c in cases], np.stack([c[2] for c in cases]),
                                                    [c[3] for c in cases], [c[4] for c in cases], 3.0, 0)
    for c, gi in zip(cases, g):
        o = oracle.fuse_candidates(c[0], c[1], c[2], c[3], c[4], 3.0, 0)
        assert np.array_equal(gi[0], o[0]) and np.array_equal(gi[1], o[1]) and gi[2] == o[2]


@pytest.mark.parametrize("seed,m,e1,e2,th", [(1, 800, 200, 300, 7.5), (2, 2000, 0, 100, 7.5), (3, 20, 5, 0, 10.0), (4, 0, 10, 10, 7.5), (5, 0, 0, 0, 7.5)])
def test_search_by_sim3(seed, m, e1, e2, th, api, oracle, synth):
    N = api.N
    sf = oracle.OrbOracle().tables()["scale_factors"]
    rng = np.random.default_rng(seed)
    kf1, kf2, p1, p2, t21, t12, lsf, keep = matchgen.sim3_case(rng, m, e1, e2, N, synth.TUM1, sf)
    g = api.DescriptorMatcher().SearchBySim3(kf1, kf2, p1, p2, t21, t12, lsf, lsf, th)
    o = oracle.search_by_sim3(kf1, kf2, p1, p2, t21, t12, lsf, lsf, th)
    assert np.array_equal(g[0], o[0]) and g[1] == o[1]
    if m >= 800:
        assert o[1] > 200


@pytest.mark.parametrize("seed,n1,n2,window", [(1, 2000, 2000, 100), (2, 1000, 1500, 30), (3, 5000, 4000, 100), (4, 0, 20, 100), (5, 20, 0, 100),
                                               (6, 300, 300, 640)])
def test_search_for_initialization(seed, n1, n2, window, api, oracle, synth):
    N = api.N
    sf = oracle.OrbOracle().tables()["scale_factors"]
    rng = np.random.default_rng(seed)
    f1, f2, prev, keep = matchgen.init_case(rng, n1, n2, N, synth.TUM1, sf)
    dm = api.DescriptorMatcher()
    for ori in (True, False):
        for ratio in (0.9, 0.7):
            g = dm.SearchForInitialization(f1, f2, prev, window, ratio, ori)
            o = oracle.search_for_initialization(f1, f2, prev, window, ratio, ori)
            assert np.array_equal(g[0], o[0]) and g[1] == o[1] and np.array_equal(g[2], o[2])
    if n1 >= 2000:
        assert o[1] > 100


@pytest.mark.parametrize("seed,m,e1,e2,nodes,only_stereo", [(1, 1000, 200, 300, 80, False), (2, 1000, 100, 100, 10, True), (3, 2500, 0, 0, 200, False),
                                                            (4, 0, 10, 10, 3, False), (5, 0, 0, 5, 3, False)])
def test_search_for_triangulation(seed, m, e1, e2, nodes, only_stereo, api, oracle, synth):
    N = api.N
    sf = oracle.OrbOracle().tables()["scale_factors"]
    rng = np.random.default_rng(seed)
    a, b, f12, cw1, tcw2, sigma2, keep = matchgen.triang_case(rng, m, e1, e2, nodes, N, synth.TUM1, sf)
    dm = api.DescriptorMatcher()
    for ori in (True, False):
        g = dm.SearchForTriangulation(a, b, f12, cw1, tcw2, synth.TUM1, sf, sigma2, only_stereo, ori)
        o = oracle.search_for_triangulation(a, b, f12, cw1, tcw2, synth.TUM1, sf, sigma2, only_stereo, ori)
        assert g[1] == o[1] and np.array_equal(g[0], o[0])
    if m >= 1000 and not only_stereo:
        assert o[1] > 100


@pytest.mark.parametrize("seed,sizes", [(1, [1, 2, 3, 5, 8, 0, 13, 40, 64, 33, 32, 31]), (2, [300, 7, 1000]), (3, list(range(0, 60))), (4, [])])
def test_distinctive_descriptors(seed, sizes, api, oracle):
    rng = np.random.default_rng(seed)
    desc, off = matchgen.distinctive_case(rng, sizes)
    if not sizes:
        desc, off = np.zeros((0, 32), np.uint8), np.zeros(1, np.int32)
    g = api.DescriptorMatcher().DistinctiveDescriptors(desc, off)
    o = oracle.distinctive_descriptors(desc, off)
    assert np.array_equal(g, o)


def test_distinctive_descriptors_many_points(api, oracle):
    """A local map's worth of map points in one launch."""
    rng = np.random.default_rng(9)
    sizes = rng.integers(2, 30, 3000).tolist()
    desc, off = matchgen.distinctive_case(rng, sizes)
    g = api.DescriptorMatcher().DistinctiveDescriptors(desc, off)
    assert np.array_equal(g, oracle.distinctive_descriptors(desc, off))
