"""CPU tests (-m "not gpu") of the oracle's F rows (Frame glue).  cv::undistortPoints is pinned
bit-exactly against python cv2 (4.13); the rest is reference-owned float arithmetic, compared with numpy restatements."""
import numpy as np
import pytest

import matchgen

f32 = np.float32


def test_undistort_pinned_to_cv2(oracle, synth):
    cv2 = pytest.importorskip("cv2")
    K = synth.TUM1
    Km = np.array([[K["fx"], 0, K["cx"]], [0, K["fy"], K["cy"]], [0, 0, 1]], np.float32)
    rng = np.random.default_rng(0)
    xy = np.stack([rng.uniform(-20, 660, 50000), rng.uniform(-20, 500, 50000)], 1).astype(np.float32)
    for D in (matchgen.TUM1_DIST, np.array([-0.28, 0.07, 1e-4, -2e-4, 0.0], np.float32), np.array([0.1, 0, 0, 0, 0], np.float32)):
        ref = cv2.undistortPoints(xy.reshape(-1, 1, 2), Km, D, None, Km).reshape(-1, 2)
        assert np.array_equal(oracle.frame_undistort_points(xy, K, D), ref)
    # k1 == 0: the reference copies the key points (Frame.cc:739-743) whatever the other coefficients are
    assert np.array_equal(oracle.frame_undistort_points(xy, K, np.array([0, 0.3, 0.01, 0, 0], np.float32)), xy)


def test_stereo_from_rgbd_and_unproject_vs_numpy(oracle, synth):
    K = synth.TUM1
    rng = np.random.default_rng(1)
    nf, rows, cols = 3, 48, 64
    depth = rng.uniform(0.3, 6, (nf, rows, cols)).astype(f32)
    depth[rng.random(depth.shape) < 0.2] = 0
    counts = [40, 0, 25]
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    xy = np.stack([rng.uniform(0, cols - 0.01, off[-1]), rng.uniform(0, rows - 0.01, off[-1])], 1).astype(f32)
    xun = (xy[:, 0] + rng.normal(0, 0.3, off[-1])).astype(f32)
    d, ur = oracle.frame_stereo_from_rgbd_batch(depth, off, xy, xun, K["bf"])
    fidx = np.repeat(np.arange(nf), counts)
    dd = depth[fidx, xy[:, 1].astype(np.int64), xy[:, 0].astype(np.int64)]
    ok = dd > 0
    assert np.array_equal(d, np.where(ok, dd, f32(-1)))
    with np.errstate(divide="ignore"):
        assert np.array_equal(ur, np.where(ok, xun - f32(K["bf"]) / dd, f32(-1)).astype(f32))
    T = np.stack([matchgen._pose(rng) for _ in range(nf)])
    rwc = np.stack([t[:3, :3].T for t in T]).astype(f32)
    ow = np.stack([matchgen._centre(t) for t in T])
    xyu = np.stack([xun, xy[:, 1]], 1)
    w, v = oracle.frame_unproject_batch(off, xyu, d, rwc, ow, K)
    assert np.array_equal(v.astype(bool), ok)
    invfx, invfy = f32(1) / f32(K["fx"]), f32(1) / f32(K["fy"])
    x = ((xyu[:, 0] - f32(K["cx"])) * d * invfx).astype(f32)
    y = ((xyu[:, 1] - f32(K["cy"])) * d * invfy).astype(f32)
    pc = np.stack([x, y, d], 1).astype(np.float64)
    ref = (np.einsum("nij,nj->ni", rwc[fidx].astype(np.float64), pc) + ow[fidx].astype(np.float64)).astype(f32)
    assert np.array_equal(w[ok], ref[ok]) and not w[~ok].any()


def test_is_in_frustum_vs_numpy(oracle, synth):
    K = synth.TUM1
    rng = np.random.default_rng(2)
    tcw, ow, Xw, normal, mi, ma, mr = matchgen.frustum_case(rng, 5, 4000, K)
    log_sf = float(f32(np.log(f32(1.2))))
    iv, px, py, pxr, lv, vc = oracle.frame_is_in_frustum_batch(tcw, ow, K, (0, 0, 640, 480), 8, log_sf, Xw, normal, mi, ma, mr, 0.5)
    assert 0.02 < iv.mean() < 0.9
    T = tcw.reshape(-1, 3, 4)
    for f in range(len(T)):
        Pc = (Xw.astype(np.float64) @ T[f, :, :3].astype(np.float64).T + T[f, :, 3].astype(np.float64)).astype(f32)
        with np.errstate(divide="ignore", invalid="ignore"):
            invz = f32(1) / Pc[:, 2]
            u = (f32(K["fx"]) * Pc[:, 0] * invz + f32(K["cx"])).astype(f32)
            v = (f32(K["fy"]) * Pc[:, 1] * invz + f32(K["cy"])).astype(f32)
            PO = Xw - ow[f]
            dist = np.sqrt((PO.astype(np.float64) ** 2).sum(1)).astype(f32)
            cosv = ((PO.astype(np.float64) * normal.astype(np.float64)).sum(1) / dist.astype(np.float64)).astype(f32)
        ok = ~(Pc[:, 2] < 0) & ~(u < 0) & ~(u > 640) & ~(v < 0) & ~(v > 480) & ~(dist < mi) & ~(dist > ma) & ~(cosv < f32(0.5))
        assert np.array_equal(iv[f].astype(bool), ok)
        assert np.array_equal(px[f][ok], u[ok]) and np.array_equal(py[f][ok], v[ok]) and np.array_equal(vc[f][ok], cosv[ok])
        assert np.array_equal(pxr[f][ok], (u - f32(K["bf"]) * invz).astype(f32)[ok])
        # PredictScale: ceil(log(max/dist) / log_sf), clamped (float log)
        lvl = np.clip(np.ceil(np.log((mr / dist).astype(f32)).astype(f32) / f32(log_sf)), 0, 7).astype(np.int32)
        assert (lv[f][ok] != lvl[ok]).mean() < 0.002      # numpy's logf may differ from glibc's in the last ulp at a ceil() edge
    s3 = Xw.astype(np.float64)
    e3 = s3 + rng.normal(0, 0.5, s3.shape)
    li = oracle.frame_lines_in_frustum_batch(tcw, s3, e3)
    for f in range(len(T)):
        zs = (s3.astype(f32).astype(np.float64) @ T[f, 2, :3].astype(np.float64) + np.float64(T[f, 2, 3])).astype(f32)
        ze = (e3.astype(f32).astype(np.float64) @ T[f, 2, :3].astype(np.float64) + np.float64(T[f, 2, 3])).astype(f32)
        assert np.array_equal(li[f].astype(bool), ~((zs < 0) & (ze < 0)))


def test_compute_stereo_matches_recovers_the_disparity(oracle, synth):
    """Frame::ComputeStereoMatches on a synthetic rectified pair with known disparity bands: the oracle's sub-pixel uRight must
    sit on the true disparity for the bulk of the matches, depth = bf / disparity, and the outlier rule must have run."""
    left, right = matchgen.stereo_pair(synth, 1000)
    ol, orr = oracle.OrbOracle(1000), oracle.OrbOracle(1000)
    kl, dl = ol.extract(left)
    kr, dr = orr.extract(right)
    K = synth.TUM1
    b = float(f32(K["bf"]) / f32(K["fx"]))
    ur, dep = oracle.frame_compute_stereo_matches(ol, orr, kl, dl, kr, dr, K["bf"], b)
    ok = ur >= 0
    assert ok.sum() > 150 and (~ok).sum() > 50
    true_d = 8.0 + 32.0 * (np.floor(kl["y"][ok] / 40) * 40) / 480 + 0.37
    err = np.abs((kl["x"][ok] - ur[ok]) - true_d)
    assert np.median(err) < 0.6 and (err < 2.0).mean() > 0.85
    assert np.array_equal(dep[ok], (f32(K["bf"]) / (kl["x"][ok] - ur[ok])).astype(f32))
    assert np.all(dep[~ok] == -1)


def test_undistort_keylines_and_grid_vs_numpy(oracle, synth):
    cv2 = pytest.importorskip("cv2")
    K = synth.TUM1
    img = synth.frame(1000, 640, 480)
    kls, _, _ = oracle.line_extract(img, 80)
    out = oracle.frame_undistort_keylines(kls, K, matchgen.TUM1_DIST, (640, 480))
    Km = np.array([[K["fx"], 0, K["cx"]], [0, K["fy"], K["cy"]], [0, 0, 1]], np.float32)
    for a, b in (("sx", "sy"), ("ex", "ey")):
        ref = cv2.undistortPoints(np.stack([kls[a], kls[b]], 1).reshape(-1, 1, 2), Km, matchgen.TUM1_DIST, None, Km).reshape(-1, 2)
        assert np.array_equal(out[a], ref[:, 0]) and np.array_equal(out[b], ref[:, 1])
    assert np.array_equal(out["sx_oct"], out["sx"]) and np.array_equal(out["class_id"], kls["class_id"]) and np.array_equal(out["octave"], kls["octave"])
    assert np.array_equal(out["pt_x"], (out["ex"] + out["sx"]) / f32(2))
    ln = np.sqrt((out["sx"] - out["ex"]).astype(np.float64) ** 2 + (out["sy"] - out["ey"]).astype(np.float64) ** 2).astype(f32)
    assert np.array_equal(out["length"], ln) and np.array_equal(out["response"], ln / f32(640))
    assert np.array_equal(oracle.frame_undistort_keylines(kls, K, np.zeros(5, f32), (640, 480)), kls)
    # grid: every feature inside the bounds sits in the cell PosInGrid gives it, ascending inside a cell
    kp, _ = oracle.OrbOracle(1000).extract(img)
    cst, idx = oracle.frame_assign_features_to_grid(kp, (0, 0, 640, 480))
    c_round = lambda v: np.where(v >= 0, np.floor(v.astype(np.float64) + 0.5), -np.floor(-v.astype(np.float64) + 0.5)).astype(np.int64)  # round(): half away from zero
    px = c_round((kp["x"] - f32(0)) * (f32(64) / f32(640)))
    py = c_round((kp["y"] - f32(0)) * (f32(48) / f32(480)))
    inside = (px >= 0) & (px < 64) & (py >= 0) & (py < 48)
    assert cst[-1] == inside.sum() and sorted(idx.tolist()) == np.nonzero(inside)[0].tolist()
    cell_of = px * 48 + py
    for c in np.unique(cell_of[inside])[:200]:
        got = idx[cst[c]:cst[c + 1]]
        assert np.array_equal(got, np.nonzero(inside & (cell_of == c))[0])
