"""CPU tests (-m "not gpu") of the oracle's G row (Frame::ComputeBoW = DBoW2 vocabulary transform): the C++ restatement against
a naive Python restatement of the DBoW2 text, the text-file loader, and the scoring / weighting variants."""
import numpy as np
import pytest

import matchgen


def _dist(a, b):
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


def naive_transform(k, L, scoring, weighting, parent, leaf, desc, weight, feats, levelsup):
    n = len(parent)
    children = [[] for _ in range(n + 1)]
    word = [0] * (n + 1)
    nw = 0
    for i in range(n):
        children[parent[i]].append(i + 1)
        if leaf[i]:
            word[i + 1] = nw
            nw += 1
    bow, fv = {}, {}
    for fi, f in enumerate(feats):
        nid_level, nid, cur, lvl = L - levelsup, 0, 0, 0
        while True:
            lvl += 1
            best, bd = None, None
            for c in children[cur]:
                d = float(_dist(f, desc[c - 1]))
                if bd is None or d < bd:
                    best, bd = c, d
            cur = best
            if lvl == nid_level:
                nid = cur
            if not children[cur]:
                break
        w = float(weight[cur - 1])
        if w > 0:
            if weighting in (0, 1):
                bow[word[cur]] = bow.get(word[cur], 0.0) + w
            elif word[cur] not in bow:
                bow[word[cur]] = w
            fv.setdefault(nid, []).append(fi)
    ids = sorted(bow)
    vals = [bow[i] for i in ids]
    must = scoring != 5
    if weighting in (0, 1) and vals and not must:
        vals = [v / float(len(vals)) for v in vals]
    if must:
        norm = 0.0
        if scoring != 1:
            for v in vals:
                norm += abs(v)
        else:
            for v in vals:
                norm += v * v
            norm = float(np.sqrt(norm))
        if norm > 0:
            vals = [v / norm for v in vals]
    return (np.asarray(ids, np.uint32), np.asarray(vals, np.float64)), {k_: fv[k_] for k_ in sorted(fv)}


@pytest.mark.parametrize("seed,k,L,scoring,weighting,n,levelsup", [(1, 10, 3, 0, 0, 400, 2), (2, 4, 5, 0, 0, 300, 4), (3, 6, 3, 1, 1, 200, 1),
                                                                    (4, 5, 3, 5, 0, 150, 2), (5, 5, 3, 0, 3, 150, 5), (6, 3, 4, 2, 2, 0, 2)])
def test_transform_vs_naive(seed, k, L, scoring, weighting, n, levelsup, oracle):
    rng = np.random.default_rng(seed)
    parent, leaf, desc, weight = matchgen.make_vocabulary(rng, k, L)
    feats = matchgen.vocabulary_features(rng, desc, leaf, n)
    v = oracle.VocOracle().create(k, L, scoring, weighting, parent, leaf, desc, weight)
    assert v.info()["n_nodes"] == len(parent) + 1 and v.info()["n_words"] == int(leaf.sum())
    (wid, wv), fv = v.transform(feats, levelsup)
    (nid_, nv), nfv = naive_transform(k, L, scoring, weighting, parent, leaf, desc, weight, feats, levelsup)
    assert np.array_equal(wid, nid_) and np.array_equal(wv, nv) and fv == nfv
    if n >= 150:
        assert len(wid) < n and any(len(x) > 1 for x in fv.values())      # words and nodes are shared by several features
        if scoring == 0:
            assert abs(wv.sum() - 1.0) < 1e-12


def test_text_file_round_trip(tmp_path, oracle):
    rng = np.random.default_rng(9)
    parent, leaf, desc, weight = matchgen.make_vocabulary(rng, 5, 3)
    path = tmp_path / "voc.txt"
    matchgen.write_vocabulary_text(path, 5, 3, 0, 0, parent, leaf, desc, weight)
    a = oracle.VocOracle().create(5, 3, 0, 0, parent, leaf, desc, weight)
    b = oracle.VocOracle()
    assert b.loadFromTextFile(path) and b.info() == a.info()
    feats = matchgen.vocabulary_features(rng, desc, leaf, 200)
    (w1, v1), f1 = a.transform(feats, 2)
    (w2, v2), f2 = b.transform(feats, 2)
    assert np.array_equal(w1, w2) and np.array_equal(v1, v2) and f1 == f2
    bad = tmp_path / "bad.txt"
    bad.write_text("99 6 0 0\n")
    assert not oracle.VocOracle().loadFromTextFile(bad)
