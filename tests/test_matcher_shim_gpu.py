"""GPU test: the matcher shims with the reference's own signatures (package host/shim/ORBmatcher.h, LineMatcher.h —
include/ORBmatcher.h:74,94 and include/LineMatcher.h:48,63 of the reference) compiled against the stand-in Frame / MapPoint /
MapLine classes and driven like Tracking.cc drives them; what they leave in mvpMapPoints / mvpMapLines is compared with the
oracle's searches on the same data.  Bar: every pointer (as an index into the map) and every count equal."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKGDIR = os.path.join(ROOT, "orb_slam2_modification_with-point-and-line-feature_b200")


def build_driver(tmp_path):
    exe = str(tmp_path / "matcher_shim_test")
    shim = os.path.join(PKGDIR, "host", "shim")
    subprocess.check_call(["g++", "-std=c++14", "-O1", os.path.join(ROOT, "tests", "cpp", "matcher_shim_test.cpp"),
                           os.path.join(shim, "ORBmatcher.cc"), os.path.join(shim, "LineMatcher.cc"), "-o", exe,
                           "-L" + PKGDIR, "-lplslam", "-Wl,-rpath," + PKGDIR])
    return exe


def noisy(desc, rng, p):
    d = desc.copy()
    d ^= np.packbits(rng.random((len(d), 256)) < p, axis=1, bitorder="little")
    return d


def make_scenario(seed, n, m, N, K, sf, oracle):
    rng = np.random.default_rng(seed)
    s = {}
    kp = np.zeros(n, N.KP_DTYPE)
    kp["x"] = rng.uniform(0, 640, n).astype(np.float32)
    kp["y"] = rng.uniform(0, 480, n).astype(np.float32)
    kp["octave"] = rng.integers(0, 8, n)
    kp["angle"] = rng.uniform(0, 360, n).astype(np.float32)
    desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    ur = np.where(rng.random(n) < 0.7, kp["x"] - rng.uniform(5, 40, n), -1).astype(np.float32)
    Tc = np.eye(4, dtype=np.float32)
    Tc[:3, 3] = rng.normal(0, 0.05, 3)
    Tl = np.eye(4, dtype=np.float32)
    P = m + 300
    src = rng.integers(0, max(n, 1), P)
    z = rng.uniform(0.5, 6, P).astype(np.float32)
    z[: P // 50] *= -1
    X = np.stack([(kp["x"][src] + rng.normal(0, 4, P) - K["cx"]) * z / K["fx"], (kp["y"][src] + rng.normal(0, 4, P) - K["cy"]) * z / K["fy"], z], 1)
    X = (X - Tc[:3, 3]).astype(np.float32)
    s["pool_pos"] = X
    s["pool_desc"] = noisy(desc[src], rng, 0.06)
    s["pool_nobs"] = rng.integers(0, 4, P).astype(np.int32)
    s["pool_bad"] = (rng.random(P) < 0.03).astype(np.uint8)
    s["pool_in_view"] = (rng.random(P) < 0.9).astype(np.uint8)
    s["pool_proj_x"] = (kp["x"][src] + rng.normal(0, 3, P)).astype(np.float32)
    s["pool_proj_y"] = (kp["y"][src] + rng.normal(0, 3, P)).astype(np.float32)
    s["pool_proj_xr"] = (s["pool_proj_x"] - 20).astype(np.float32)
    s["pool_level"] = np.clip(kp["octave"][src] + rng.integers(0, 2, P), 0, 7).astype(np.int32)
    s["pool_view_cos"] = rng.uniform(0.99, 1.0, P).astype(np.float32)
    perm = rng.permutation(P)
    s["last_mp"] = np.where(rng.random(m) < 0.9, perm[:m], -1).astype(np.int32)
    s["last_outlier"] = (rng.random(m) < 0.05).astype(np.uint8)
    lk = np.zeros(m, N.KP_DTYPE)
    lk["octave"] = kp["octave"][src[np.clip(s["last_mp"], 0, None)]]
    lku = lk.copy()
    lku["angle"] = rng.uniform(0, 360, m).astype(np.float32)
    s["last_keys"], s["last_keys_un"] = lk, lku
    s["cur_keys_un"], s["cur_desc"], s["cur_u_right"] = kp, desc, ur
    s["cur_mp"] = np.where(rng.random(n) < 0.05, rng.integers(0, P, n), -1).astype(np.int32)
    s["cur_tcw"], s["last_tcw"] = Tc, Tl
    s["K"] = np.array([K["fx"], K["fy"], K["cx"], K["cy"], K["bf"], np.float32(K["bf"]) / np.float32(K["fx"]), 0, 0, 640, 480], np.float32)
    s["scale_factors"] = np.asarray(sf, np.float32)
    s["img_size"] = np.array([640, 480], np.int32)
    s["local_points"] = rng.permutation(P)[: int(0.8 * P)].astype(np.int32)
    # ---- lines: 3-D segments whose projections lie (mostly) in the image; the current frame sees noisy copies of some ----
    PL, ML = 90, 60
    uv = np.stack([rng.uniform(-50, 690, (PL, 2)), rng.uniform(-40, 520, (PL, 2))], 2)   # [line, endpoint, (u, v)]
    zz = rng.uniform(1, 5, (PL, 2))
    zz[:6, 0] *= -1
    zz[6:9] *= -1
    R, t = Tc[:3, :3].astype(np.float64), Tc[:3, 3].astype(np.float64)
    pts = np.stack([(uv[..., 0] - K["cx"]) * zz / K["fx"], (uv[..., 1] - K["cy"]) * zz / K["fy"], zz], 2)   # camera frame
    pw = (pts - t) @ R
    s["lpool_start"], s["lpool_end"] = np.ascontiguousarray(pw[:, 0]), np.ascontiguousarray(pw[:, 1])
    s["lpool_desc"] = rng.integers(0, 256, (PL, 32), dtype=np.uint8)
    s["lpool_nobs"] = rng.integers(0, 3, PL).astype(np.int32)
    s["lpool_bad"] = (rng.random(PL) < 0.05).astype(np.uint8)
    s["lpool_in_view"] = (rng.random(PL) < 0.85).astype(np.uint8)
    lperm = rng.permutation(PL)
    s["last_ml"] = np.where(rng.random(ML) < 0.85, lperm[:ML], -1).astype(np.int32)
    s["last_line_outlier"] = (rng.random(ML) < 0.08).astype(np.uint8)
    lkl = np.zeros(ML, N.KL_DTYPE)
    lkl["class_id"] = np.arange(ML)
    lkl["octave"] = 0
    s["last_kl_un"] = lkl
    s["local_lines"] = rng.permutation(PL)[:70].astype(np.int32)
    # current lines: the oracle's projection of every pool line, jittered
    pk, pidx = oracle.project_lines(s["lpool_start"], s["lpool_end"], np.zeros(PL, N.KL_DTYPE), np.ones(PL, np.uint8), Tc[:3].reshape(-1), K,
                                    (0, 0, 640, 480), (640, 480))
    take = rng.permutation(len(pk))[: min(len(pk), 45)]
    cur = pk[take].copy()
    for fld in ("sx", "ex"):
        cur[fld] += rng.normal(0, 1.5, len(cur)).astype(np.float32)
    cur["pt_x"], cur["pt_y"] = (cur["sx"] + cur["ex"]) / 2, (cur["sy"] + cur["ey"]) / 2
    cur["length"] = np.hypot(cur["sx"] - cur["ex"], cur["sy"] - cur["ey"]).astype(np.float32)
    cur["angle"] = np.arctan2(cur["ey"] - cur["sy"], cur["ex"] - cur["sx"]).astype(np.float32)
    s["cur_kl_un"] = cur
    s["cur_ldesc"] = noisy(s["lpool_desc"][pidx[take]], rng, 0.04)
    s["cur_ml"] = np.where(rng.random(len(cur)) < 0.1, rng.integers(0, PL, len(cur)), -1).astype(np.int32)
    return s


def expected(s, th3, th2, N, K, sf, oracle):
    keep = []
    nobs, bad = s["pool_nobs"], s["pool_bad"]
    cur_mp = s["cur_mp"].copy()
    bounds = (0, 0, 640, 480)

    def frame_view(mp):
        claimed = ((mp >= 0) & (nobs[np.clip(mp, 0, None)] > 0)).astype(np.int32)
        return N.make_frame_view(s["cur_keys_un"], s["cur_desc"], s["cur_u_right"], claimed, bounds, K, s["cur_tcw"][:3].reshape(-1), sf, keep)

    out = {}
    # C3
    lm = s["last_mp"]
    li = np.clip(lm, 0, None)
    valid = (lm >= 0) & (s["last_outlier"] == 0)
    lv = N.make_lastframe_view(valid, s["pool_pos"][li], s["pool_desc"][li], s["last_keys"]["octave"], s["last_keys_un"]["angle"], nobs[li] > 0,
                               s["last_tcw"][:3].reshape(-1), keep)
    m3, n3 = oracle.search_last_frame(frame_view(cur_mp), lv, th3, False, True)[:2]
    had = cur_mp >= 0
    cur_mp = np.where(m3 >= 0, lm[np.clip(m3, 0, None)], np.where(had, cur_mp, -1)).astype(np.int32)
    out["c3"] = (n3, cur_mp.copy())
    # D3
    lnobs, lbad = s["lpool_nobs"], s["lpool_bad"]
    cur_ml = s["cur_ml"].copy()

    def line_frame_view(ml):
        claimed = ((ml >= 0) & (lnobs[np.clip(ml, 0, None)] > 0)).astype(np.uint8)
        return N.make_lineframe_view(s["cur_kl_un"], s["cur_ldesc"], claimed, s["cur_tcw"][:3].reshape(-1), K, bounds, (640, 480), keep)

    lml = s["last_ml"]
    lli = np.clip(lml, 0, None)
    lvalid = (lml >= 0) & (s["last_line_outlier"] == 0) & (lbad[lli] == 0)
    mlv = N.make_mapline_view(s["lpool_start"][lli], s["lpool_end"][lli], s["last_kl_un"], s["lpool_desc"][lli], lvalid, keep)
    r3, nl3, rel3, _ = oracle.line_search_by_projection(line_frame_view(cur_ml), mlv)
    if rel3:
        cur_ml[:] = -1
    cur_ml = np.where(r3 >= 0, lml[np.clip(r3, 0, None)], cur_ml).astype(np.int32)
    out["d3"] = (nl3, cur_ml.copy())
    # C2
    loc = s["local_points"]
    mv = N.make_mappoint_view(s["pool_desc"][loc], (s["pool_in_view"][loc] != 0) & (bad[loc] == 0), s["pool_proj_x"][loc], s["pool_proj_y"][loc],
                              s["pool_proj_xr"][loc], s["pool_level"][loc], s["pool_view_cos"][loc], nobs[loc] > 0, keep)
    m2, n2 = oracle.search_local_points(frame_view(cur_mp), mv, th2, 0.8)[:2]
    cur_mp = np.where(m2 >= 0, loc[np.clip(m2, 0, None)], cur_mp).astype(np.int32)
    out["c2"] = (n2, cur_mp.copy())
    # D5
    ll = s["local_lines"]
    mlv5 = N.make_mapline_view(s["lpool_start"][ll], s["lpool_end"][ll], np.zeros(len(ll), N.KL_DTYPE), s["lpool_desc"][ll],
                               (s["lpool_in_view"][ll] != 0) & (lbad[ll] == 0), keep)
    r5, nl5, rel5, _ = oracle.line_search_by_projection(line_frame_view(cur_ml), mlv5)
    if rel5:
        cur_ml[:] = -1
    cur_ml = np.where(r5 >= 0, ll[np.clip(r5, 0, None)], cur_ml).astype(np.int32)
    out["d5"] = (nl5, cur_ml.copy())
    return out


@pytest.mark.parametrize("seed,n,m,th3", [(101, 1000, 900, 15.0), (102, 700, 1200, 30.0), (103, 40, 30, 7.0)])
def test_shims_with_reference_signatures(seed, n, m, th3, api, oracle, synth, tmp_path):
    N = api.N
    K = synth.TUM1
    sf = oracle.OrbOracle().tables()["scale_factors"]
    exe = build_driver(tmp_path)
    s = make_scenario(seed, n, m, N, K, sf, oracle)
    d = tmp_path / "scn"
    d.mkdir()
    for k, v in s.items():
        np.ascontiguousarray(v).tofile(str(d / (k + ".bin")))
    lines = subprocess.run([exe, str(d), str(th3), "3"], capture_output=True, text=True, check=True).stdout.splitlines()
    got = {}
    for ln in lines:
        f = ln.split()
        got[f[0]] = (int(f[1]), np.array([int(x) for x in f[2:]], np.int32))
    want = expected(s, th3, 3.0, N, K, sf, oracle)
    for k in ("c3", "d3", "c2", "d5"):
        assert got[k][0] == want[k][0], (k, got[k][0], want[k][0])
        assert np.array_equal(got[k][1], want[k][1]), k
    if n >= 700:
        assert want["c3"][0] > 40 and want["c2"][0] > 200 and want["d3"][0] + want["d5"][0] > 5
