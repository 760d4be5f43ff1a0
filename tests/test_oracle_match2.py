"""CPU tests: the oracle's C4-C7 / D6 functions against an independent, deliberately naive Python restatement of the
reference loops (ORBmatcher.cc:247-410, 729-872; LineMatcher.cpp:489-525, 1174-1204, 1296-1330; KeyFrame.cc:773-797) and
hand-made cases.  The reference cannot be built here (OpenCV/DBoW2/Pangolin), so these rows are pinned by restatement only."""
import numpy as np
import pytest

import matchgen


def ham(a, b):
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


def py_bow(A, B, mode, nn_ratio, check_ori):
    """A, B: dicts with desc, angle, valid (or None), fv (dict node -> list)."""
    n_out = len(B["desc"]) if mode == 0 else len(A["desc"])
    match = [-1] * n_out
    matchedB = [False] * len(B["desc"])
    hist = [[] for _ in range(30)]
    nm = 0
    for node in sorted(set(A["fv"]) & set(B["fv"])):
        for ia in A["fv"][node]:
            if A["valid"] is not None and not A["valid"][ia]:
                continue
            b1, b2, bi = 256, 256, -1
            for ib in B["fv"][node]:
                if matchedB[ib] or (mode == 1 and B["valid"] is not None and not B["valid"][ib]):
                    continue
                d = ham(A["desc"][ia], B["desc"][ib])
                if d < b1:
                    b2, b1, bi = b1, d, ib
                elif d < b2:
                    b2 = d
            if (b1 <= 50 if mode == 0 else b1 < 50) and np.float32(b1) < np.float32(nn_ratio) * np.float32(b2):
                matchedB[bi] = True
                slot = bi if mode == 0 else ia
                match[slot] = ia if mode == 0 else bi
                if check_ori:
                    rot = np.float32(A["angle"][ia]) - np.float32(B["angle"][bi])
                    if rot < 0:
                        rot = np.float32(rot + np.float32(360))
                    v = float(np.float32(rot * np.float32(30 / 360.0)))
                    b = int(np.floor(v + 0.5))          # round(): half away from zero, v >= 0
                    hist[0 if b == 30 else b].append(slot)
                nm += 1
    if check_ori:
        sizes = [len(h) for h in hist]
        order = []
        m1 = m2 = m3 = 0
        i1 = i2 = i3 = -1
        for i, s in enumerate(sizes):
            if s > m1:
                m3, m2, m1, i3, i2, i1 = m2, m1, s, i2, i1, i
            elif s > m2:
                m3, m2, i3, i2 = m2, s, i2, i
            elif s > m3:
                m3, i3 = s, i
        if m2 < 0.1 * m1:
            i2 = i3 = -1
        elif m3 < 0.1 * m1:
            i3 = -1
        for i in range(30):
            if i not in (i1, i2, i3):
                for slot in hist[i]:
                    match[slot] = -1
                    nm -= 1
    return np.asarray(match, np.int32), nm


def _unpack_bow(v, keepers):
    import ctypes as C
    def arr(addr, n, dt):
        if not addr or n == 0:
            return np.zeros(0, dt)
        return np.ctypeslib.as_array(C.cast(addr, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), (n,)).copy()
    desc = arr(v.desc, v.n * 32, np.uint8).reshape(-1, 32)
    angle = arr(v.angle, v.n, np.float32)
    valid = arr(v.valid, v.n, np.uint8) if v.valid else None
    ids = arr(v.node_id, v.n_nodes, np.uint32)
    off = arr(v.node_off, v.n_nodes + 1, np.int32)
    fi = arr(v.feat_idx, int(off[-1]) if v.n_nodes else 0, np.uint32)
    fv = {int(ids[k]): [int(x) for x in fi[off[k]:off[k + 1]]] for k in range(v.n_nodes)}
    return dict(desc=desc, angle=angle, valid=valid, fv=fv)


@pytest.mark.parametrize("seed,nA,nB,nodes,mode", [(1, 300, 300, 40, 0), (2, 300, 300, 40, 1), (3, 60, 500, 5, 0), (4, 500, 60, 5, 1), (5, 0, 50, 4, 0),
                                                   (6, 50, 0, 4, 1)])
def test_bow_oracle_vs_python(seed, nA, nB, nodes, mode, oracle, pkg):
    N = pkg.load_native() if hasattr(pkg, "load_native") else __import__("importlib").import_module(pkg.__name__ + "._native")
    rng = np.random.default_rng(seed)
    a, b, keep = matchgen.bow_case(rng, nA, nB, nodes, N, mode)
    A, B = _unpack_bow(a, keep), _unpack_bow(b, keep)
    for ori in (True, False):
        o = oracle.search_bow(a, b, mode, 0.75, ori)
        p = py_bow(A, B, mode, 0.75, ori)
        assert np.array_equal(o[0], p[0]) and o[1] == p[1]
    if nA >= 300 and nB >= 300:
        assert o[1] > 50


def test_knn_ratio_and_fuse_and_mad(oracle):
    rng = np.random.default_rng(7)
    cur = rng.integers(0, 256, (80, 32), dtype=np.uint8)
    ref = matchgen.noisy(cur[rng.integers(0, 80, 60)], rng, 0.05)
    ref[:5] = rng.integers(0, 256, (5, 32), dtype=np.uint8)       # unrelated lines: ratio test fails
    idx, dist = oracle.hamming_knn2(ref, cur)
    # ratio rule
    exp = np.full(80, -1, np.int32)
    cnt = 0
    for i in range(60):
        if np.float32(dist[i, 0]) / np.float32(dist[i, 1]) < 0.75:
            exp[idx[i, 0]] = i
            cnt += 1
    m, n = oracle.line_match_knn_ratio(ref, cur)
    assert np.array_equal(m, exp) and n == cnt and cnt > 40
    assert oracle.line_match_knn_ratio(ref, cur[:1])[1] == 0      # one train row: the reference reads out of bounds
    # MAD rule
    d0, d1 = dist[:, 0].astype(np.float64), dist[:, 1].astype(np.float64)
    e = np.sort(d1 - d0)
    med12 = e[len(e) // 2]
    mad12 = 1.4826 * np.sort(np.abs(d1 - d0 - med12))[len(e) // 2]
    med = np.sort(d0)[len(d0) // 2]
    mad = 1.4826 * np.sort(np.abs(d0 - med))[len(d0) // 2]
    pairs, m1, m2 = oracle.line_search_for_triangulation(ref, cur)
    assert m1 == mad and m2 == mad12
    exp_pairs = [(i, idx[i, 0]) for i in range(60) if d1[i] - d0[i] > 0.1 * mad12]
    assert [tuple(p) for p in pairs] == exp_pairs and len(exp_pairs) > 10
    # Fuse rule: dist < 1.5 * min(100, dist)  <=>  0 < dist < 150
    ml = np.concatenate([cur[:3], matchgen.noisy(cur[3:6], rng, 0.02), rng.integers(0, 256, (6, 32), dtype=np.uint8)])
    valid = np.ones(len(ml), np.uint8)
    valid[4] = 0
    tdx, nf = oracle.line_fuse_candidates(ml, valid, cur)
    i1, dd = oracle.hamming_knn2(ml, cur)
    exp_t = [int(i1[i, 0]) if (valid[i] and 0 < dd[i, 0] < 150) else -1 for i in range(len(ml))]
    assert list(tdx) == exp_t and nf == sum(t >= 0 for t in exp_t)
    assert tdx[0] == -1 and tdx[3] == 3                              # identical descriptor (dist 0) never fuses


def test_pose_searches_smoke_and_invariants(oracle, pkg, synth):
    """C4 / C5 oracle: invariants that do not depend on the implementation (claimed features never matched, matches point to
    valid points whose descriptor distance is within the threshold, behind-camera points never match in C5)."""
    import importlib
    N = importlib.import_module(pkg.__name__ + "._native")
    sf = oracle.OrbOracle().tables()["scale_factors"]
    for mode in (2, 3):
        rng = np.random.default_rng(20 + mode)
        fv, pv, ow, log_sf, keep = matchgen.pose_case(rng, 1200, 900, N, synth.TUM1, sf, mode)
        if mode == 2:
            match, n = oracle.search_keyframe_points(fv, pv, ow, log_sf, 10.0, 100, False)
            th_d = 100
        else:
            match, n = oracle.search_sim3_points(fv, pv, ow, log_sf, 10)
            th_d = 50
        assert n == int((match >= 0).sum()) and n > 100
        claimed, fdesc, valid, pdesc = keep[3], keep[1], keep[5], keep[7]
        for i2 in np.nonzero(match >= 0)[0]:
            assert not claimed[i2] and valid[match[i2]]
            assert ham(fdesc[i2], pdesc[match[i2]]) <= th_d
        assert len(set(match[match >= 0])) == n                     # a map point is assigned at most once
