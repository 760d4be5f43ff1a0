"""GPU parity tests (-m gpu) for the F rows (Frame glue) through the C ABI vs the oracle (and cv2 for
cv::undistortPoints).  Bar: bit-exact, floats included."""
import numpy as np
import pytest

import matchgen

pytestmark = pytest.mark.gpu
f32 = np.float32


@pytest.mark.parametrize("n", [0, 1, 1000, 200000])
def test_undistort_points(n, api, oracle, synth):
    rng = np.random.default_rng(n)
    xy = np.stack([rng.uniform(-20, 660, n), rng.uniform(-20, 500, n)], 1).astype(f32)
    dm = api.DescriptorMatcher()
    for D in (matchgen.TUM1_DIST, np.array([-0.28, 0.07, 1e-4, -2e-4, 0.0], f32), np.array([0, 0.3, 0, 0, 0], f32)):
        g = dm.UndistortPoints(xy, synth.TUM1, D)
        assert np.array_equal(g, oracle.frame_undistort_points(xy, synth.TUM1, D))
    if n:
        cv2 = pytest.importorskip("cv2")
        K = synth.TUM1
        Km = np.array([[K["fx"], 0, K["cx"]], [0, K["fy"], K["cy"]], [0, 0, 1]], f32)
        ref = cv2.undistortPoints(xy.reshape(-1, 1, 2), Km, matchgen.TUM1_DIST, None, Km).reshape(-1, 2)
        assert np.array_equal(dm.UndistortPoints(xy, K, matchgen.TUM1_DIST), ref)


@pytest.mark.parametrize("counts,device_depth", [([1000, 0, 997, 1003], False), ([1000] * 8, True), ([0, 0], False), ([], False)])
def test_stereo_from_rgbd_and_unproject(counts, device_depth, api, oracle, synth):
    K = synth.TUM1
    rng = np.random.default_rng(len(counts))
    nf, rows, cols = len(counts), 480, 640
    depth = rng.uniform(0.3, 6, (max(nf, 1), rows, cols)).astype(f32)[:nf]
    depth[rng.random(depth.shape) < 0.2] = 0
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    tot = int(off[-1])
    xy = np.stack([rng.uniform(0, cols - 0.01, tot), rng.uniform(0, rows - 0.01, tot)], 1).astype(f32)
    xun = (xy[:, 0] + rng.normal(0, 0.3, tot)).astype(f32)
    dm = api.DescriptorMatcher()
    if nf == 0:
        return
    if device_depth:
        import torch
        dd = torch.from_numpy(depth).cuda()
        g = dm.StereoFromRGBDBatch((nf, rows, cols), off, xy, xun, K["bf"], depth_dev_ptr=dd.data_ptr())
    else:
        g = dm.StereoFromRGBDBatch(depth, off, xy, xun, K["bf"])
    o = oracle.frame_stereo_from_rgbd_batch(depth, off, xy, xun, K["bf"])
    assert np.array_equal(g[0], o[0]) and np.array_equal(g[1], o[1])
    T = np.stack([matchgen._pose(rng) for _ in range(nf)])
    rwc = np.stack([t[:3, :3].T for t in T]).astype(f32)
    ow = np.stack([matchgen._centre(t) for t in T])
    xyu = np.stack([xun, xy[:, 1]], 1)
    gw, gv = dm.UnprojectBatch(off, xyu, o[0], rwc, ow, K)
    ow_, ov = oracle.frame_unproject_batch(off, xyu, o[0], rwc, ow, K)
    assert np.array_equal(gw, ow_) and np.array_equal(gv, ov)
    if tot:
        assert 0.5 < gv.mean() < 0.95


def test_stereo_from_rgbd_rejects_positions_outside_the_image(api, synth):
    dm = api.DescriptorMatcher()
    depth = np.ones((1, 48, 64), f32)
    with pytest.raises(api.N.PlError):
        dm.StereoFromRGBDBatch(depth, [0, 1], np.array([[64.0, 3.0]], f32), np.array([64.0], f32), 40.0)


@pytest.mark.parametrize("seed,nf,m", [(1, 10, 3000), (2, 1, 1), (3, 300, 3000), (4, 0, 10), (5, 4, 0)])
def test_is_in_frustum(seed, nf, m, api, oracle, synth):
    K = synth.TUM1
    rng = np.random.default_rng(seed)
    tcw, ow, Xw, normal, mi, ma, mr = matchgen.frustum_case(rng, nf, m, K)
    log_sf = float(f32(np.log(f32(1.2))))
    dm = api.DescriptorMatcher()
    g = dm.IsInFrustumBatch(tcw, ow, K, (0, 0, 640, 480), 8, log_sf, Xw, normal, mi, ma, mr, 0.5)
    o = oracle.frame_is_in_frustum_batch(tcw, ow, K, (0, 0, 640, 480), 8, log_sf, Xw, normal, mi, ma, mr, 0.5)
    for a, b in zip(g, o):
        assert np.array_equal(a, b)
    if nf >= 10 and m >= 3000:
        assert 0.02 < o[0].mean() < 0.9
    s3 = Xw.astype(np.float64)
    e3 = s3 + rng.normal(0, 0.5, s3.shape)
    assert np.array_equal(dm.LinesInFrustumBatch(tcw, s3, e3), oracle.frame_lines_in_frustum_batch(tcw, s3, e3))


def test_frustum_feeds_local_point_search(api, oracle, synth):
    """F4 -> C2 chained: the planes IsInFrustum writes are the map-point view SearchByProjection reads."""
    N = api.N
    K = synth.TUM1
    sf = oracle.OrbOracle().tables()["scale_factors"]
    rng = np.random.default_rng(11)
    fv, pv, ow, log_sf, inv_s2, keep = matchgen.fuse_case(rng, 1200, 300, N, K, sf)
    import ctypes as C
    m = pv.n
    arr = lambda addr, dt, cnt: np.frombuffer((C.c_char * (cnt * np.dtype(dt).itemsize)).from_address(addr), dt).copy()
    Xw, nrm = arr(pv.world_pos, f32, 3 * m).reshape(m, 3), arr(pv.normal, f32, 3 * m).reshape(m, 3)
    mi, ma, mr = arr(pv.min_dist_inv, f32, m), arr(pv.max_dist_inv, f32, m), arr(pv.max_dist, f32, m)
    desc = arr(pv.desc, np.uint8, 32 * m).reshape(m, 32)
    tcw = np.array(list(fv.tcw), f32)[None]
    dm = api.DescriptorMatcher()
    iv, px, py, pxr, lv, vc = dm.IsInFrustumBatch(tcw, ow[None], K, (0, 0, 640, 480), 8, log_sf, Xw, nrm, mi, ma, mr, 0.5)
    mv = N.make_mappoint_view(desc, iv[0], px[0], py[0], pxr[0], lv[0], vc[0], None, keep)
    g = dm.SearchByProjectionLocalPoints(fv, mv, 3.0, 0.8)
    o = oracle.search_local_points(fv, mv, 3.0, 0.8)
    assert np.array_equal(g[0], o[0]) and g[1] == o[1] and o[1] > 200


@pytest.mark.parametrize("seed,nfeat,size", [(1000, 1000, (640, 480)), (1001, 2000, (640, 480)), (3000, 2000, (1241, 376))])
def test_compute_stereo_matches(seed, nfeat, size, api, oracle, synth):
    """F5: row-band Hamming search + SAD refinement on the device-resident pyramids of both extractors, vs the oracle."""
    w, h = size
    left, right = matchgen.stereo_pair(synth, seed, w, h)
    el = api.ORBextractor(nfeat, 1.2, 8, 20, 7, max_cols=w, max_rows=h)
    er = api.ORBextractor(nfeat, 1.2, 8, 20, 7, max_cols=w, max_rows=h)
    kl, dl = el(left)
    kr, dr = er(right)
    ol, orr = oracle.OrbOracle(nfeat), oracle.OrbOracle(nfeat)
    okl, odl = ol.extract(left)
    okr, odr = orr.extract(right)
    assert np.array_equal(dl, odl) and np.array_equal(dr, odr)
    K = synth.TUM1
    b = float(f32(K["bf"]) / f32(K["fx"]))
    g = api.DescriptorMatcher().ComputeStereoMatches(el, er, kl, dl, kr, dr, K["bf"], b)
    o = oracle.frame_compute_stereo_matches(ol, orr, okl, odl, okr, odr, K["bf"], b)
    assert np.array_equal(g[0], o[0]) and np.array_equal(g[1], o[1])
    assert (o[0] >= 0).sum() > 150


def test_compute_stereo_matches_degenerate(api, oracle, synth):
    """No right key points / identical images: every SAD score is 0, so the median rule (thDist = 0, `first < thDist` never
    true) drops every match — in the reference as well."""
    img = synth.frame(1000, 640, 480)
    el, er = api.ORBextractor(1000, 1.2, 8, 20, 7), api.ORBextractor(1000, 1.2, 8, 20, 7)
    kl, dl = el(img)
    kr, dr = er(img)
    ol, orr = oracle.OrbOracle(1000), oracle.OrbOracle(1000)
    ol.extract(img)
    orr.extract(img)
    dm = api.DescriptorMatcher()
    K = synth.TUM1
    b = float(f32(K["bf"]) / f32(K["fx"]))
    g = dm.ComputeStereoMatches(el, er, kl, dl, kr, dr, K["bf"], b)
    o = oracle.frame_compute_stereo_matches(ol, orr, kl, dl, kr, dr, K["bf"], b)
    assert np.array_equal(g[0], o[0]) and np.array_equal(g[1], o[1]) and np.all(o[0] == -1)
    g = dm.ComputeStereoMatches(el, er, kl, dl, kr[:0], dr[:0], K["bf"], b)
    assert np.all(g[0] == -1) and np.all(g[1] == -1)


@pytest.mark.parametrize("n", [0, 1, 80, 5000])
def test_undistort_keylines(n, api, oracle, synth):
    rng = np.random.default_rng(n)
    kls = np.zeros(n, api.N.KL_DTYPE)
    for f_ in ("sx", "ex"):
        kls[f_] = rng.uniform(0, 640, n).astype(f32)
    for f_ in ("sy", "ey"):
        kls[f_] = rng.uniform(0, 480, n).astype(f32)
    kls["class_id"] = np.arange(n)
    kls["octave"] = rng.integers(0, 2, n)
    dm = api.DescriptorMatcher()
    for D in (matchgen.TUM1_DIST, np.zeros(5, f32)):
        g = dm.UndistortKeyLines(kls, synth.TUM1, D, (640, 480))
        o = oracle.frame_undistort_keylines(kls, synth.TUM1, D, (640, 480))
        for name in api.N.KL_DTYPE.names:
            if name == "angle":     # atan2f: CUDA libm vs glibc, tolerance of the line path
                assert np.allclose(g[name], o[name], rtol=1e-4, atol=1e-6)
            else:
                assert np.array_equal(g[name], o[name]), name


@pytest.mark.parametrize("n,bounds", [(0, (0, 0, 640, 480)), (1, (0, 0, 640, 480)), (1000, (0, 0, 640, 480)), (3000, (-12.5, -9.25, 655.0, 490.5)),
                                      (2000, (100, 100, 300, 200))])
def test_assign_features_to_grid(n, bounds, api, oracle):
    rng = np.random.default_rng(n)
    kp, _, _ = matchgen.rand_frame(rng, n, api.N)
    g = api.DescriptorMatcher().AssignFeaturesToGrid(kp, bounds)
    o = oracle.frame_assign_features_to_grid(kp, bounds)
    assert np.array_equal(g[0], o[0]) and np.array_equal(g[1], o[1])
