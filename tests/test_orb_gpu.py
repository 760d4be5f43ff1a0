"""GPU parity tests (-m gpu): the CUDA ORB path, called through the C ABI, against the CPU oracle on the same
seeded inputs and against the committed golden fixtures.  Bar: bit-exact keypoints (incl. float angle, order,
octave), descriptors, pyramid planes, blurred planes and FAST candidate lists."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def kp_matrix(k):
    return np.stack([k["x"], k["y"], k["size"], k["angle"], k["response"], k["octave"].astype(np.float32)], 1)


def _flat():
    flat = np.full((240, 320), 128, np.uint8)
    flat[100:140, 150:200] = 200
    return flat


CASES = [("cfgA_seed1000", lambda s: s.frame(1000, 640, 480), 1000),
         ("small_seed7", lambda s: s.frame(7, 320, 240), 500),
         ("kitti_seed3000", lambda s: s.frame(3000, 1241, 376), 2000),
         ("sparse", lambda s: _flat(), 500)]


@pytest.mark.parametrize("name,make,nf", CASES, ids=[c[0] for c in CASES])
def test_extract_matches_oracle_and_golden(name, make, nf, synth, oracle, api, golden_dir):
    img = make(synth)
    ex = api.ORBextractor(nf, 1.2, 8, 20, 7, max_cols=img.shape[1], max_rows=img.shape[0], max_batch=1)
    k, d = ex(img)
    o = oracle.OrbOracle(nf)
    ok, od = o.extract(img)
    # intermediates first: they localise a failure
    for l in range(8):
        assert np.array_equal(ex.pyramid_level(l, bordered=True), o.level_bordered(l)), f"pyramid level {l}"
    for l in range(8):
        cx, cy, cr = ex.candidates(l)
        ox, oy, orr = o.level_candidates(l)
        assert len(cx) == len(ox), f"candidate count level {l}: {len(cx)} vs {len(ox)}"
        assert np.array_equal(cx, ox) and np.array_equal(cy, oy) and np.array_equal(cr, orr), f"candidates level {l}"
    for l in range(8):
        ob = o.level_blurred(l)
        if ob is not None:
            assert np.array_equal(ex.blurred_level(l), ob), f"blurred level {l}"
    assert len(k) == len(ok)
    assert np.array_equal(kp_matrix(k), kp_matrix(ok))
    assert (k["class_id"] == -1).all()
    assert np.array_equal(d, od)
    g = np.load(os.path.join(golden_dir, f"orb_{name}.npz"))
    assert np.array_equal(kp_matrix(k), g["keypoints"]) and np.array_equal(d, g["descriptors"])
    ex.close()


def test_getters_match_reference_tables(api, oracle):
    ex = api.ORBextractor(1000, 1.2, 8, 20, 7)
    t = oracle.OrbOracle(1000).tables()
    assert ex.GetLevels() == 8 and abs(ex.GetScaleFactor() - np.float32(1.2)) < 1e-12
    assert np.array_equal(ex.GetScaleFactors(), t["scale_factors"])
    assert np.array_equal(ex.GetInverseScaleFactors(), t["inv_scale_factors"])
    assert np.array_equal(ex.GetScaleSigmaSquares(), t["sigma2"])
    assert np.array_equal(ex.GetInverseScaleSigmaSquares(), t["inv_sigma2"])
    assert np.array_equal(ex.features_per_level(), t["per_level"])


def test_empty_and_bad_input(api):
    ex = api.ORBextractor(500, 1.2, 8, 20, 7)
    k, d = ex(np.zeros((0, 0), np.uint8))
    assert len(k) == 0 and d.shape == (0, 32)
    with pytest.raises(AssertionError):
        ex(np.zeros((480, 640), np.float32))
    with pytest.raises(api.PlError):
        ex(np.zeros((60, 60), np.uint8))  # level 7 would be smaller than the FAST border


def test_strided_input_and_size_change(api, synth, oracle):
    big = synth.frame(21, 700, 500)
    view = big[10:490, 30:670]  # 640x480 view with row stride 700
    ex = api.ORBextractor(1000, 1.2, 8, 20, 7, max_cols=700, max_rows=500)
    k, d = ex(view)
    ok, od = oracle.OrbOracle(1000).extract(np.ascontiguousarray(view))
    assert np.array_equal(kp_matrix(k), kp_matrix(ok)) and np.array_equal(d, od)
    k2, d2 = ex(big)  # same handle, different size
    ok2, od2 = oracle.OrbOracle(1000).extract(big)
    assert np.array_equal(kp_matrix(k2), kp_matrix(ok2)) and np.array_equal(d2, od2)


def test_batch_equals_single(api, synth, oracle):
    frames = synth.frames(6000, 20)  # 20 frames, chunk size 8 -> 3 chunks incl. a ragged one
    ex = api.ORBextractor(1000, 1.2, 8, 20, 7, max_batch=8)
    kps, desc, cnt = ex.extract_batch(frames)
    o = oracle.OrbOracle(1000)
    for i in (0, 1, 7, 8, 15, 19):
        ok, od = o.extract(frames[i])
        assert cnt[i] == len(ok)
        assert np.array_equal(kp_matrix(kps[i, :cnt[i]]), kp_matrix(ok)), f"frame {i}"
        assert np.array_equal(desc[i, :cnt[i]], od), f"frame {i}"
    # idempotence: a second pass over the same batch gives identical bytes
    kps2, desc2, cnt2 = ex.extract_batch(frames)
    assert np.array_equal(cnt, cnt2) and np.array_equal(desc, desc2) and np.array_equal(kps, kps2)


@pytest.mark.parametrize("params", [(300, 1.2, 8, 20, 7), (1500, 1.5, 5, 30, 10), (800, 2.0, 4, 20, 7), (1000, 1.1, 12, 12, 5)])
def test_other_extractor_parameters(params, api, synth, oracle):
    img = synth.frame(77, 640, 480)
    nf, sf, nl, it, mt = params
    ex = api.ORBextractor(nf, sf, nl, it, mt)
    k, d = ex(img)
    o = oracle.OrbOracle(nf, sf, nl, it, mt)
    ok, od = o.extract(img)
    for l in range(nl):
        assert np.array_equal(ex.pyramid_level(l, bordered=True), o.level_bordered(l)), f"pyramid level {l}"
    assert np.array_equal(kp_matrix(k), kp_matrix(ok)) and np.array_equal(d, od)
