"""CPU tests: Hamming oracle vs numpy, vs the cv2.BFMatcher golden fixtures, and (if importable) vs cv2 itself."""
import os

import numpy as np
import pytest


def np_hamming(a, b):
    return np.unpackbits(a ^ b, axis=-1).sum(-1).astype(np.int32)


def test_pairs_vs_numpy(oracle):
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, (500, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (500, 32), dtype=np.uint8)
    b[:10] = a[:10]
    b[10:20] = ~a[10:20]
    d = oracle.hamming_pairs(a, b)
    assert np.array_equal(d, np_hamming(a, b))
    assert d[:10].tolist() == [0] * 10 and d[10:20].tolist() == [256] * 10


@pytest.mark.parametrize("n", [1024, 2048])
def test_knn2_golden(n, synth, oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, f"knn_{n}.npz"))
    q, t = synth.descriptor_sets(n)
    idx, dist = oracle.hamming_knn2(q, t, threads=4)
    assert np.array_equal(idx, g["idx"]) and np.array_equal(dist, g["dist"])
    idx1, dist1 = oracle.hamming_knn2(q[:64], t, threads=1)
    assert np.array_equal(idx1, g["idx"][:64]) and np.array_equal(dist1, g["dist"][:64])


def test_knn2_ties_and_short_train(oracle):
    t = np.zeros((3, 32), np.uint8)
    q = np.zeros((2, 32), np.uint8)
    idx, dist = oracle.hamming_knn2(q, t)
    assert idx.tolist() == [[0, 1], [0, 1]] and dist.tolist() == [[0, 0], [0, 0]]   # ties -> lowest train index
    idx, dist = oracle.hamming_knn2(q, t[:1])
    assert idx.tolist() == [[0, -1], [0, -1]] and dist.tolist() == [[0, -1], [0, -1]]


def test_candidates(oracle):
    rng = np.random.default_rng(2)
    q = rng.integers(0, 256, (10, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (40, 32), dtype=np.uint8)
    lens = rng.integers(0, 9, 10)
    off = np.r_[0, np.cumsum(lens)].astype(np.int32)
    ci = rng.integers(0, 40, off[-1]).astype(np.int32)
    d = oracle.hamming_candidates(q, t, off, ci)
    ref = np.concatenate([np_hamming(q[i][None], t[ci[off[i]:off[i + 1]]]) for i in range(10)] + [np.zeros(0, np.int32)])
    assert np.array_equal(d, ref)
