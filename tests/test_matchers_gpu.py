"""GPU parity tests (-m gpu): projection searches of ORBmatcher and LineMatcher through the C ABI vs the oracle.
Bar: every match index and count bit-exact."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PKG = "orb_slam2_modification_with-point-and-line-feature_b200"


@pytest.fixture(scope="module")
def seq(synth):
    gray, depth, T = synth.room_sequence(14)
    return gray, depth, T


@pytest.fixture(scope="module")
def feats(seq, oracle):
    gray = seq[0]
    ob = oracle.OracleBackend()
    return ob.extract_orb(gray), ob.extract_lines(gray)


class Recorder:
    """Wraps a backend and records every matcher call's outputs."""

    NAMES = ("search_last_frame", "search_local_points", "search_last_frame_batch", "search_local_points_batch", "line_search_batch",
             "search_local_map_batch")

    def __init__(self, b):
        self.b = b
        self.log = []

    def __getattr__(self, k):
        f = getattr(self.b, k)
        if k in self.NAMES:
            def w(*a):
                r = f(*a)
                self.log.append((k, r))
                return r
            return w
        return f


def _flatten(log):
    """Calls of the log as (kind, result), grouped by kind: the point searches and the line searches may run on several host
    threads, so their relative order in the log is not defined."""
    out = []
    for k, r in log:
        if k == "search_local_map_batch":   # IsInFrustum + C2 in one call: its C2 part is compared with the frame-by-frame C2 calls
            out += [("search_local_points", x[:2]) for x in r]
        elif k.endswith("_batch"):
            out += [(k[:-6], x) for x in r]
        else:
            out.append((k, r))
    # the two line searches of a frame (D3, D5) may also run side by side: inside a kind the results are ordered by content,
    # so both arms are compared as multisets of (kind, result)
    return sorted(out, key=lambda e: (e[0], len(e[1][0]), e[1][0].tobytes(), tuple(int(x) for x in e[1][1:])))


def test_sequence_matchers_bit_exact(seq, feats, api, oracle, pkg):
    fe = importlib.import_module(PKG + ".frontend")
    gray, depth, T = seq
    ob = Recorder(oracle.OracleBackend())
    sf = ob.b.scale_factors()
    so = fe.TrackingFrontEnd(ob).run(gray, depth, T, sf, features=feats, batch=False)
    ref = _flatten(ob.log)
    assert len(ref) > 30
    for batch in (True, False):
        gb = Recorder(fe.GpuBackend(api, 480, 640))
        sg = fe.TrackingFrontEnd(gb).run(gray, depth, T, sf, features=feats, batch=batch)
        assert sg == so, f"batch={batch}"
        got = _flatten(gb.log)
        assert len(got) == len(ref)
        n_matches = 0
        for (kg, rg), (ko, ro) in zip(got, ref):
            assert kg == ko
            assert np.array_equal(rg[0], ro[0]) and tuple(rg[1:]) == tuple(ro[1:]), kg    # every index and count bit-exact
            n_matches += rg[1]
        assert n_matches > 1000   # the schedule really matched things
    assert any(r.get("c2_matches", 0) > 50 for r in so) and any(r.get("d3_matches", 0) > 5 for r in so)
    assert any(r.get("d5_matches", 0) > 5 for r in so)


def test_sequence_with_device_glue_bit_exact(seq, feats, api, oracle, pkg):
    """The same schedule with the Frame glue (UnprojectStereo, IsInFrustum) on the device (F rows): both arms still agree
    bit for bit, and the searches still match."""
    fe = importlib.import_module(PKG + ".frontend")
    gray, depth, T = seq
    ob = Recorder(oracle.OracleBackend())
    sf = ob.b.scale_factors()
    so = fe.TrackingFrontEnd(ob, device_glue=True).run(gray, depth, T, sf, features=feats, batch=False)
    ref = _flatten([x for x in ob.log if x[0].startswith(("search", "line_search"))])
    gb = Recorder(fe.GpuBackend(api, 480, 640))
    sg = fe.TrackingFrontEnd(gb, device_glue=True).run(gray, depth, T, sf, features=feats, batch=True)
    assert sg == so
    got = _flatten([x for x in gb.log if x[0].startswith(("search", "line_search"))])
    assert len(got) == len(ref) > 30
    for (kg, rg), (ko, ro) in zip(got, ref):
        assert kg == ko and np.array_equal(rg[0], ro[0]) and tuple(rg[1:]) == tuple(ro[1:]), kg
    assert sum(r.get("c2_matches", 0) for r in so) > 200 and sum(r.get("c3_matches", 0) for r in so) > 1000


def test_search_local_map_equals_frustum_plus_search(api, oracle, synth):
    """pl_orb_search_local_map_batch (Tracking::SearchLocalPoints: IsInFrustum + C2 with the projections left on the device) against
    the oracle's two steps and against the library's own two calls: snapshots shared by runs of frames, an empty snapshot, an empty
    frame, a snapshot used by two separate runs."""
    import matchgen
    N = api.N
    rng = np.random.default_rng(77)
    sf = oracle.OrbOracle().tables()["scale_factors"]
    K = synth.TUM1
    log_sf = float(np.float32(np.log(np.float32(sf[1] / sf[0]))))
    keep, snaps, raw = [], [], []
    for m in (2500, 0, 900):
        tcw, ow, Xw, nrm, mi, ma, mr = matchgen.frustum_case(rng, 4, m, K)
        desc = rng.integers(0, 256, (m, 32), dtype=np.uint8)
        snaps.append(N.make_localmap_view(Xw, nrm, desc, mi, ma, mr, None, keep))
        raw.append((tcw, ow, Xw, nrm, mi, ma, mr, desc))
    mof = [0, 0, 0, 1, 2, 2, 0, 0]
    nfeat = [800, 1000, 0, 500, 700, 300, 1200, 64]
    fvs, ows, tcws = [], [], []
    for i, (k, nf) in enumerate(zip(mof, nfeat)):
        tcw, ow = raw[k][0][i % 4], raw[k][1][i % 4]
        kp, desc, ur = matchgen.rand_frame(rng, nf, N)
        m = snaps[k].n
        if m and nf:   # features on top of projected map points, with their descriptors (so that there is something to match)
            iv, px, py, pxr, lv, vc = oracle.frame_is_in_frustum_batch(tcw[None], ow[None], K, (0, 0, 640, 480), 8, log_sf, *raw[k][2:7])
            vis = np.flatnonzero(iv[0])[: nf // 2]
            kp["x"][: len(vis)] = px[0][vis] + rng.normal(0, 0.7, len(vis)).astype(np.float32)
            kp["y"][: len(vis)] = py[0][vis] + rng.normal(0, 0.7, len(vis)).astype(np.float32)
            kp["octave"][: len(vis)] = lv[0][vis]
            desc[: len(vis)] = matchgen.noisy(raw[k][7][vis], rng, 0.06)
        claimed = (rng.random(nf) < 0.1).astype(np.int32)
        fvs.append(N.make_frame_view(kp, desc, ur, claimed, (0, 0, 640, 480), K, tcw, sf, keep))
        ows.append(ow)
        tcws.append(tcw)
    ows = np.stack(ows)
    m = api.DescriptorMatcher()
    got = m.SearchLocalMapBatch(fvs, ows, snaps, mof, 0.5, log_sf, 3.0, 0.8)
    total = 0
    for i, (gm, gn, giv) in enumerate(got):
        om, on, oiv = oracle.search_local_map(fvs[i], ows[i], snaps[mof[i]], 0.5, log_sf, 3.0, 0.8)
        assert np.array_equal(gm, om) and gn == on and giv == oiv, i
        total += gn
    assert total > 500
    # the library's own two calls give the same
    for i in (0, 4, 7):
        k = mof[i]
        if snaps[k].n == 0:
            continue
        iv, px, py, pxr, lv, vc = m.IsInFrustumBatch(tcws[i][None], ows[i][None], K, (0, 0, 640, 480), 8, log_sf, *raw[k][2:7], 0.5)
        mv = N.make_mappoint_view(raw[k][7], iv[0], px[0], py[0], pxr[0], lv[0], vc[0], None, keep)
        (tm, tn), = m.SearchByProjectionLocalPointsBatch([fvs[i]], [mv], 3.0, 0.8)
        assert np.array_equal(tm, got[i][0]) and tn == got[i][1] and int(iv[0].sum()) == got[i][2]
    assert m.SearchLocalMapBatch([], np.zeros((0, 3), np.float32), snaps, [], 0.5, log_sf, 3.0, 0.8) == []


def test_batch_with_ragged_and_empty_instances(api, oracle, synth):
    """Batched searches with instances of different sizes, including empty frames / empty point sets."""
    N = api.N
    rng = np.random.default_rng(5)
    sf = oracle.OrbOracle().tables()["scale_factors"]
    K = synth.TUM1
    keep, fvs, mvs, lvs = [], [], [], []
    for n, m in [(500, 700), (0, 10), (40, 0), (1200, 3000), (3, 3)]:
        kp, desc, ur = _rand_frame(rng, n, N)
        fvs.append(N.make_frame_view(kp, desc, ur, (rng.random(n) < 0.1).astype(np.int32), (0, 0, 640, 480), K,
                                     np.eye(4, dtype=np.float32)[:3].reshape(-1), sf, keep))
        src = rng.integers(0, max(n, 1), m)
        md = (desc[src] if n else rng.integers(0, 256, (m, 32), dtype=np.uint8)).copy()
        md ^= np.packbits(rng.random((m, 256)) < 0.05, axis=1, bitorder="little")
        px = (kp["x"][src] if n else np.zeros(m)) + rng.normal(0, 3, m)
        py = (kp["y"][src] if n else np.zeros(m)) + rng.normal(0, 3, m)
        lvl = np.clip((kp["octave"][src] if n else np.zeros(m, np.int64)) + rng.integers(0, 2, m), 0, 7)
        mvs.append(N.make_mappoint_view(md, rng.random(m) < 0.9, px, py, px - 20, lvl, rng.uniform(0.99, 1.0, m), rng.random(m) < 0.5, keep))
        z = rng.uniform(0.5, 6, m).astype(np.float32)
        X = np.stack([(px - K["cx"]) * z / K["fx"], (py - K["cy"]) * z / K["fy"], z], 1).astype(np.float32)
        lvs.append(N.make_lastframe_view(rng.random(m) < 0.95, X, md, lvl, rng.uniform(0, 360, m), rng.random(m) < 0.6,
                                         np.eye(4, dtype=np.float32)[:3].reshape(-1), keep))
    m = api.DescriptorMatcher()
    g2 = m.SearchByProjectionLocalPointsBatch(fvs, mvs, 3.0, 0.8)
    g3 = m.SearchByProjectionLastFrameBatch(fvs, lvs, 15.0)
    for i in range(len(fvs)):
        o2 = oracle.search_local_points(fvs[i], mvs[i], 3.0, 0.8)
        o3 = oracle.search_last_frame(fvs[i], lvs[i], 15.0)
        assert np.array_equal(g2[i][0], o2[0]) and g2[i][1] == o2[1], i
        assert np.array_equal(g3[i][0], o3[0]) and g3[i][1] == o3[1], i
    assert g2[3][1] > 100 and g3[3][1] > 100


def _rand_frame(rng, n, N):
    kp = np.zeros(n, N.KP_DTYPE)
    kp["x"] = rng.uniform(0, 640, n).astype(np.float32)
    kp["y"] = rng.uniform(0, 480, n).astype(np.float32)
    kp["octave"] = rng.integers(0, 8, n)
    kp["angle"] = rng.uniform(0, 360, n).astype(np.float32)
    desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    ur = np.where(rng.random(n) < 0.7, kp["x"] - rng.uniform(5, 40, n), -1).astype(np.float32)
    return kp, desc, ur


@pytest.mark.parametrize("seed,n,m,th", [(1, 1000, 2500, 3.0), (2, 50, 10, 1.0), (3, 2000, 300, 5.0), (4, 0, 20, 3.0), (5, 300, 0, 3.0)])
def test_local_points_random(seed, n, m, th, api, oracle, synth):
    """Random (dense, colliding) inputs: many points compete for the same features -> the claim order matters."""
    N = api.N
    rng = np.random.default_rng(seed)
    kp, desc, ur = _rand_frame(rng, n, N)
    sf = oracle.OrbOracle().tables()["scale_factors"]
    K = synth.TUM1
    claimed = (rng.random(n) < 0.1).astype(np.int32)
    keep = []
    fv = N.make_frame_view(kp, desc, ur, claimed, (0, 0, 640, 480), K, np.eye(4, dtype=np.float32)[:3].reshape(-1), sf, keep)
    src = rng.integers(0, max(n, 1), m)
    mdesc = (desc[src] if n else rng.integers(0, 256, (m, 32), dtype=np.uint8)).copy()
    flips = rng.random((m, 256)) < 0.05
    mdesc ^= np.packbits(flips, axis=1, bitorder="little")
    px = (kp["x"][src] if n else np.zeros(m)) + rng.normal(0, 3, m)
    py = (kp["y"][src] if n else np.zeros(m)) + rng.normal(0, 3, m)
    lvl = np.clip((kp["octave"][src] if n else np.zeros(m, np.int64)) + rng.integers(0, 2, m), 0, 7)
    mv = N.make_mappoint_view(mdesc, rng.random(m) < 0.9, px, py, px - 20, lvl, rng.uniform(0.99, 1.0, m), rng.random(m) < 0.5, keep)
    g = api.DescriptorMatcher().SearchByProjectionLocalPoints(fv, mv, th, 0.8)
    o = oracle.search_local_points(fv, mv, th, 0.8)
    assert np.array_equal(g[0], o[0]) and g[1] == o[1]
    if n >= 1000 and m >= 1000:
        assert g[1] > 200


@pytest.mark.parametrize("seed,n,m,th", [(11, 1000, 1000, 15.0), (12, 1000, 1000, 30.0), (13, 64, 900, 7.0), (14, 1500, 0, 15.0)])
def test_last_frame_random(seed, n, m, th, api, oracle, synth):
    N = api.N
    rng = np.random.default_rng(seed)
    kp, desc, ur = _rand_frame(rng, n, N)
    sf = oracle.OrbOracle().tables()["scale_factors"]
    K = synth.TUM1
    keep = []
    T = np.eye(4, dtype=np.float32)
    T[:3, 3] = rng.normal(0, 0.05, 3)
    fv = N.make_frame_view(kp, desc, ur, (rng.random(n) < 0.05).astype(np.int32), (0, 0, 640, 480), K, T[:3].reshape(-1), sf, keep)
    src = rng.integers(0, n, m)
    z = rng.uniform(0.5, 6, m).astype(np.float32)
    z[: m // 50] *= -1                                             # some points behind the camera
    X = np.stack([(kp["x"][src] + rng.normal(0, 4, m) - K["cx"]) * z / K["fx"], (kp["y"][src] + rng.normal(0, 4, m) - K["cy"]) * z / K["fy"], z], 1)
    X = (X - T[:3, 3]).astype(np.float32)
    ldesc = desc[src].copy()
    ldesc ^= np.packbits(rng.random((m, 256)) < 0.06, axis=1, bitorder="little")
    Tl = np.eye(4, dtype=np.float32)
    Tl[2, 3] = 0.3 if seed == 12 else 0.0                          # seed 12: camera moved forward -> bForward window
    lv = N.make_lastframe_view(rng.random(m) < 0.95, X, ldesc, kp["octave"][src], rng.uniform(0, 360, m), rng.random(m) < 0.6,
                               Tl[:3].reshape(-1), keep)
    for ori in (True, False):
        g = api.DescriptorMatcher().SearchByProjectionLastFrame(fv, lv, th, False, ori)
        o = oracle.search_last_frame(fv, lv, th, False, ori)
        assert np.array_equal(g[0], o[0]) and g[1] == o[1]
    if m:
        assert o[1] > 20


def test_line_project_and_match_cases(api, oracle, synth):
    N = api.N
    K = synth.TUM1
    rng = np.random.default_rng(21)
    n = 200
    kl = np.zeros(n, N.KL_DTYPE)
    kl["class_id"] = np.arange(n)
    z1, z2 = rng.uniform(-1, 5, n), rng.uniform(-1, 5, n)         # endpoints in front of / behind the camera
    s3 = np.stack([rng.uniform(-3, 3, n), rng.uniform(-2, 2, n), z1], 1)
    e3 = np.stack([rng.uniform(-3, 3, n), rng.uniform(-2, 2, n), z2], 1)
    e3[:10, 1] = s3[:10, 1]                                        # horizontal in 3-D
    e3[10:20, 0] = s3[10:20, 0]                                    # vertical
    T = np.eye(4, dtype=np.float32)
    T[:3, 3] = (0.1, -0.05, 0.2)
    m = api.DescriptorMatcher()
    valid = np.random.default_rng(22).random(n) < 0.9
    g = m.project_lines(s3, e3, kl, valid, T[:3].reshape(-1), K, (0, 0, 640, 480), (640, 480))
    o = oracle.project_lines(s3, e3, kl, valid, T[:3].reshape(-1), K, (0, 0, 640, 480), (640, 480))
    assert np.array_equal(g[1], o[1]) and len(g[1]) > 20
    for fld in ("sx", "sy", "ex", "ey", "num_pixels"):
        assert np.array_equal(g[0][fld], o[0][fld]), fld
    # match: crafted current lines = noisy copies of the projected ones
    pk = o[0]
    cur = pk.copy()[: len(pk) // 2]
    for fld in ("sx", "ex"):
        cur[fld] += rng.normal(0, 2, len(cur)).astype(np.float32)
    pd = rng.integers(0, 256, (len(pk), 32), dtype=np.uint8)
    cd = pd[: len(cur)].copy()
    cd ^= np.packbits(rng.random((len(cur), 256)) < 0.05, axis=1, bitorder="little")
    claimed = (rng.random(len(cur)) < 0.2).astype(np.uint8)
    for cl in (None, claimed):
        gm = m.match_lines(pk, pd, cur, cd, cl)
        om = oracle.match_lines(pk, pd, cur, cd, cl)
        assert np.array_equal(gm[0], om[0]) and gm[1:] == om[1:]
    # relaxed retry: random descriptors -> < 20 % matches in the strict pass
    cd2 = rng.integers(0, 256, cd.shape, dtype=np.uint8)
    gm = m.match_lines(pk, pd, cur, cd2, claimed)
    om = oracle.match_lines(pk, pd, cur, cd2, claimed)
    assert om[2] == 1 and np.array_equal(gm[0], om[0]) and gm[1:] == om[1:]
    # empty sides
    assert m.match_lines(pk[:0], pd[:0], cur, cd, None)[1] == 0
    assert len(m.project_lines(s3[:0], e3[:0], kl[:0], valid[:0], T[:3].reshape(-1), K, (0, 0, 640, 480), (640, 480))[0]) == 0
