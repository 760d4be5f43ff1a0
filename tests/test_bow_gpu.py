"""GPU parity tests (-m gpu) for the G row (Frame::ComputeBoW = DBoW2 vocabulary transform) through the
C ABI vs the oracle.  Bar: word ids, node ids, feature lists and every double of the BowVector bit-exact."""
import numpy as np
import pytest

import matchgen

pytestmark = pytest.mark.gpu


def _same(g, o):
    (gw, gv), gf = g
    (ow, ov), of = o
    return np.array_equal(gw, ow) and np.array_equal(gv, ov) and gf == of


@pytest.mark.parametrize("seed,k,L,scoring,weighting,n,levelsup", [(1, 10, 4, 0, 0, 1000, 2), (2, 10, 3, 0, 0, 2000, 4), (3, 6, 3, 1, 1, 500, 1),
                                                                    (4, 5, 3, 5, 0, 300, 2), (5, 5, 3, 0, 3, 300, 5), (6, 3, 4, 2, 2, 0, 2),
                                                                    (7, 20, 2, 3, 0, 8192, 1), (8, 2, 9, 4, 2, 777, 4)])
def test_transform(seed, k, L, scoring, weighting, n, levelsup, api, oracle):
    rng = np.random.default_rng(seed)
    parent, leaf, desc, weight = matchgen.make_vocabulary(rng, k, L)
    feats = matchgen.vocabulary_features(rng, desc, leaf, n)
    v = api.ORBVocabulary().create(k, L, scoring, weighting, parent, leaf, desc, weight)
    o = oracle.VocOracle().create(k, L, scoring, weighting, parent, leaf, desc, weight)
    assert v.info() == o.info()
    assert _same(v.transform(feats, levelsup), o.transform(feats, levelsup))


def test_transform_batch_ragged_and_text_file(tmp_path, api, oracle):
    rng = np.random.default_rng(21)
    parent, leaf, desc, weight = matchgen.make_vocabulary(rng, 10, 3)
    path = tmp_path / "voc.txt"
    matchgen.write_vocabulary_text(path, 10, 3, 0, 0, parent, leaf, desc, weight)
    v = api.ORBVocabulary()
    assert v.loadFromTextFile(path)
    o = oracle.VocOracle()
    assert o.loadFromTextFile(path) and v.info() == o.info()
    sets = [matchgen.vocabulary_features(rng, desc, leaf, n) for n in (1000, 0, 1, 37, 2000, 1000)]
    g = v.transform_batch(sets, 2)
    for gi, s in zip(g, sets):
        assert _same(gi, o.transform(s, 2))
    bad = tmp_path / "bad.txt"
    bad.write_text("99 6 0 0\n")
    assert not api.ORBVocabulary().loadFromTextFile(bad)
    assert not api.ORBVocabulary().loadFromTextFile(tmp_path / "missing.txt")


def test_transform_feeds_search_by_bow(api, oracle, synth):
    """extract -> ComputeBoW -> SearchByBoW chained through the C ABI: the FeatureVector comes out in pl_bow_view's layout."""
    N = api.N
    rng = np.random.default_rng(5)
    ex = api.ORBextractor(1000, 1.2, 8, 20, 7)
    ka, da = ex(synth.frame(1000, 640, 480))
    kb, db = ex(np.roll(synth.frame(1000, 640, 480), 3, axis=1))
    # a vocabulary grown around the descriptors of the two frames
    parent, leaf, desc, weight = matchgen.make_vocabulary(rng, 10, 3, leaf_depth_jitter=False, stop_frac=0.0)
    v = api.ORBVocabulary().create(10, 3, 0, 0, parent, leaf, desc, weight)
    o = oracle.VocOracle().create(10, 3, 0, 0, parent, leaf, desc, weight)
    (ga, gb) = v.transform_batch([da, db], 2)
    assert _same(ga, o.transform(da, 2)) and _same(gb, o.transform(db, 2))
    keep = []
    va = N.make_bow_view(ka["angle"], da, None, ga[1], keep)
    vb = N.make_bow_view(kb["angle"], db, None, gb[1], keep)
    g = api.DescriptorMatcher().SearchByBoWBatch([va], [vb], 0, 0.9, True)[0]
    r = oracle.search_bow(va, vb, 0, 0.9, True)
    assert np.array_equal(g[0], r[0]) and g[1] == r[1] and r[1] > 50
