"""CPU tests (-m "not gpu"): the ORB oracle against the committed golden fixtures (generated with real cv2 calls by
tests/golden/make_golden.py) and, when cv2 is importable, against the OpenCV primitives directly."""
import os

import numpy as np
import pytest


def _flat():
    flat = np.full((240, 320), 128, np.uint8)
    flat[100:140, 150:200] = 200
    return flat


CASES = [("cfgA_seed1000", lambda s: s.frame(1000, 640, 480), 1000),
         ("small_seed7", lambda s: s.frame(7, 320, 240), 500),
         ("kitti_seed3000", lambda s: s.frame(3000, 1241, 376), 2000),
         ("sparse", lambda s: _flat(), 500)]


def kp_matrix(k):
    return np.stack([k["x"], k["y"], k["size"], k["angle"], k["response"], k["octave"].astype(np.float32)], 1)


@pytest.mark.parametrize("name,make,nf", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_golden(name, make, nf, synth, oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, f"orb_{name}.npz"))
    o = oracle.OrbOracle(nf)
    k, d = o.extract(make(synth))
    assert np.array_equal(kp_matrix(k), g["keypoints"])        # bit-exact incl. angle and quadtree order
    assert np.array_equal(d, g["descriptors"])
    assert [len(o.level_candidates(l)[0]) for l in range(8)] == g["cand_counts"].tolist()
    assert (k["class_id"] == -1).all()


def test_tables(oracle):
    t = oracle.OrbOracle(1000).tables()
    assert t["per_level"].tolist() == [217, 181, 151, 126, 105, 87, 73, 60]
    assert t["umax"].tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    t = oracle.OrbOracle(2000).tables()
    assert t["per_level"].tolist() == [434, 362, 302, 251, 209, 175, 145, 122]


def test_empty_image(oracle):
    k, d = oracle.OrbOracle(100).extract(np.zeros((0, 0), np.uint8).reshape(0, 1)[:, :1].reshape(0, 1))
    assert len(k) == 0 and len(d) == 0


def test_octree_small_cases(oracle):
    # single key, duplicate-free tiny sets, N larger than the number of keys
    x, y, r = oracle.distribute_octtree([5.0], [7.0], [33.0], 16, 624, 16, 464, 10)
    assert (x.tolist(), y.tolist(), r.tolist()) == ([5.0], [7.0], [33.0])
    xs = np.array([1, 300, 301, 600, 10, 11], np.float32)
    ys = np.array([1, 200, 201, 440, 400, 401], np.float32)
    rs = np.array([10, 20, 30, 40, 50, 60], np.float32)
    # two tight pairs never separate: a round without growth ends the loop (ORBextractor.cc:669), best response kept
    x, y, r = oracle.distribute_octtree(xs, ys, rs, 16, 624, 16, 464, 100)
    assert (x.tolist(), r.tolist()) == ([11.0, 301.0, 1.0, 600.0], [60.0, 30.0, 10.0, 40.0])
    x, y, r = oracle.distribute_octtree(xs, ys, rs, 16, 624, 16, 464, 1)
    assert len(x) >= 1


try:
    import cv2  # noqa: F811
    cv2.setNumThreads(1)
except Exception:  # pragma: no cover
    cv2 = None

needs_cv2 = pytest.mark.skipif(cv2 is None, reason="python cv2 not importable")


@needs_cv2
@pytest.mark.parametrize("seed,w,h", [(1000, 640, 480), (11, 333, 217), (12, 97, 61)])
def test_primitives_vs_cv2(seed, w, h, synth, oracle):
    img = synth.frame(seed, w, h)
    for (dw, dh) in [(int(round(w / 1.2)), int(round(h / 1.2))), (w // 2, h // 2), (w - 1, h - 3)]:
        assert np.array_equal(cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR), oracle.resize_linear(img, dw, dh))
    assert np.array_equal(cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101), oracle.border_reflect101(img, 19))
    assert np.array_equal(cv2.GaussianBlur(img, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101),
                          oracle.gaussian_blur7(img))
    for th in (20, 7):
        for nms in (True, False):
            fd = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=nms, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
            ref = np.array([(p.pt[0], p.pt[1], p.response if nms else 0) for p in fd.detect(img)], np.float32).reshape(-1, 3)
            xs, ys, rs = oracle.fast_detect(img, th, nms)
            mine = np.stack([xs, ys, rs if nms else np.zeros_like(rs)], 1)
            assert np.array_equal(ref, mine)


@needs_cv2
def test_fast_atan2_vs_cv2(oracle):
    rng = np.random.default_rng(0)
    y = rng.normal(0, 1000, 5000).astype(np.float32)
    x = rng.normal(0, 1000, 5000).astype(np.float32)
    y[:50] = 0
    x[25:75] = 0
    ref = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(y, x)], np.float32)
    assert np.array_equal(ref, oracle.fast_atan2(y, x))
