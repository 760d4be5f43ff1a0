"""GPU parity tests (-m gpu): Hamming kernels through the C ABI vs the oracle / golden cv2.BFMatcher vectors."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def np_hamming(a, b):
    return np.unpackbits(a ^ b, axis=-1).sum(-1).astype(np.int32)


def test_pairs(api, oracle):
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, (3001, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (3001, 32), dtype=np.uint8)
    b[:7] = a[:7]
    m = api.DescriptorMatcher()
    assert np.array_equal(m.DescriptorDistance(a, b), oracle.hamming_pairs(a, b))


@pytest.mark.parametrize("n", [1024, 2048])
def test_knn2_golden(n, api, synth, golden_dir):
    g = np.load(os.path.join(golden_dir, f"knn_{n}.npz"))
    q, t = synth.descriptor_sets(n)
    idx, dist = api.DescriptorMatcher().knnMatch2(q, t)
    assert np.array_equal(idx, g["idx"]) and np.array_equal(dist, g["dist"])


@pytest.mark.parametrize("nq,nt", [(1, 1), (3, 2), (80, 80), (1000, 777), (129, 4097), (4096, 4096)])
def test_knn2_vs_oracle(nq, nt, api, oracle):
    rng = np.random.default_rng(nq * 7 + nt)
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    if nt > 4:
        t[nt // 2] = t[0]           # duplicate rows: tie -> lowest index
        q[0] = t[0]
    idx, dist = api.DescriptorMatcher().knnMatch2(q, t)
    oi, od = oracle.hamming_knn2(q, t, threads=8)
    assert np.array_equal(idx, oi) and np.array_equal(dist, od)


def test_knn2_full_size_properties(api, synth):
    """16k x 16k (config 4): self-match property — querying the train set returns itself at distance 0 — and
    agreement of the best distance with a numpy check on a sample."""
    n = 16384
    q, t = synth.descriptor_sets(n)
    m = api.DescriptorMatcher()
    idx, dist = m.knnMatch2(t, t)
    first_of = {}
    for i, row in enumerate(map(bytes, t)):
        first_of.setdefault(row, i)
    expect = np.array([first_of[bytes(r)] for r in t], np.int32)
    assert np.array_equal(idx[:, 0], expect) and (dist[:, 0] == 0).all()
    idx, dist = m.knnMatch2(q, t)
    for i in range(0, n, 997):
        d = np_hamming(q[i][None], t)
        order = np.lexsort((np.arange(n), d))
        assert idx[i].tolist() == order[:2].tolist() and dist[i].tolist() == d[order[:2]].tolist()


@pytest.mark.parametrize("n", [8192, 16384])
def test_knn2_8k_16k_equal_cv2_bfmatcher_in_full(n, api, synth):
    """Config 4 at its upper sizes, every row: indices and distances of cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2) — what the
    reference's LineMatcher calls (src/LineMatcher.cpp:497-503) — for the ORB-like and the LBD-like descriptor sets."""
    cv2 = pytest.importorskip("cv2")
    for seed_shift in (0, 1):
        q, t = synth.descriptor_sets(n) if seed_shift == 0 else synth.descriptor_sets(n)[::-1]
        idx, dist = api.DescriptorMatcher().knnMatch2(q, t)
        m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)
        ci = np.array([[a.trainIdx, b.trainIdx] for a, b in m], np.int32)
        cd = np.array([[a.distance, b.distance] for a, b in m]).astype(np.int32)
        assert np.array_equal(dist, cd)
        # equal distances: cv2 keeps the first index it met, the library the lowest index: compare indices where the row is free of ties
        same = idx == ci
        if not same.all():
            bad = np.nonzero(~same.all(axis=1))[0]
            for i in bad:
                d = np.unpackbits(q[i][None] ^ t, axis=1).sum(1)
                assert (d == dist[i, 0]).sum() > 1 or (d == dist[i, 1]).sum() > 1, f"row {i}: indices differ without a tie"


def test_candidates(api, oracle):
    rng = np.random.default_rng(5)
    q = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (1000, 32), dtype=np.uint8)
    lens = rng.integers(0, 70, 300)
    lens[5] = 0
    off = np.r_[0, np.cumsum(lens)].astype(np.int32)
    ci = rng.integers(0, 1000, off[-1]).astype(np.int32)
    d = api.DescriptorMatcher().candidate_distances(q, t, off, ci)
    assert np.array_equal(d, oracle.hamming_candidates(q, t, off, ci))
