"""Synthetic inputs for the C4-C7 / D6 matcher tests (shared by the CPU and GPU suites)."""
import numpy as np


def rand_frame(rng, n, N):
    kp = np.zeros(n, N.KP_DTYPE)
    kp["x"] = rng.uniform(0, 640, n).astype(np.float32)
    kp["y"] = rng.uniform(0, 480, n).astype(np.float32)
    kp["octave"] = rng.integers(0, 8, n)
    kp["angle"] = rng.uniform(0, 360, n).astype(np.float32)
    desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    ur = np.where(rng.random(n) < 0.7, kp["x"] - rng.uniform(5, 40, n), -1).astype(np.float32)
    return kp, desc, ur


def noisy(desc, rng, p):
    d = desc.copy()
    d ^= np.packbits(rng.random((len(d), 256)) < p, axis=1, bitorder="little")
    return d


def pose_case(rng, n, m, N, K, sf, mode, claimed_frac=0.05):
    """A frame with n features and m map points that project near them.  mode 2 = C4, 3 = C5.
    Returns (frame_view, posepoint_view, ow, log_sf, keep)."""
    keep = []
    kp, desc, ur = rand_frame(rng, n, N)
    T = np.eye(4, dtype=np.float32)
    ang = rng.normal(0, 0.03, 3)
    Rx = np.array([[1, 0, 0], [0, np.cos(ang[0]), -np.sin(ang[0])], [0, np.sin(ang[0]), np.cos(ang[0])]])
    Ry = np.array([[np.cos(ang[1]), 0, np.sin(ang[1])], [0, 1, 0], [-np.sin(ang[1]), 0, np.cos(ang[1])]])
    T[:3, :3] = (Rx @ Ry).astype(np.float32)
    T[:3, 3] = rng.normal(0, 0.05, 3)
    fv = N.make_frame_view(kp, desc, ur, (rng.random(n) < claimed_frac).astype(np.int32), (0, 0, 640, 480), K, T[:3].reshape(-1), sf, keep)
    src = rng.integers(0, max(n, 1), m)
    z = rng.uniform(0.5, 8, m)
    if m:
        z[: max(m // 40, 1)] *= -1                         # behind the camera
    xs = (kp["x"][src] if n else np.zeros(m)) + rng.normal(0, 4, m)
    ys = (kp["y"][src] if n else np.zeros(m)) + rng.normal(0, 4, m)
    Xc = np.stack([(xs - K["cx"]) * z / K["fx"], (ys - K["cy"]) * z / K["fy"], z], 1)
    R, t = T[:3, :3].astype(np.float64), T[:3, 3].astype(np.float64)
    Xw = ((Xc - t) @ R).astype(np.float32)                  # R^T (Xc - t)
    # Ow = -R^T t evaluated like the cv::Mat expression (double accumulation, rounded once)
    ow = np.array([np.float32(-(sum(np.float64(T[k, r]) * np.float64(T[k, 3]) for k in range(3)))) for r in range(3)], np.float32)
    dist = np.linalg.norm(Xw.astype(np.float64) - ow.astype(np.float64), axis=1)
    lvl = rng.integers(0, 8, m)
    log_sf = np.float32(np.log(np.float32(1.2)))
    max_raw = (dist * np.float32(1.2) ** (lvl + rng.uniform(-0.4, 0.4, m))).astype(np.float32)   # PredictScale ~ lvl
    max_inv = (np.float32(1.2) * max_raw).astype(np.float32)
    min_inv = (np.float32(0.8) * max_raw / np.float32(1.2) ** 7).astype(np.float32)
    if m:
        min_inv[:: 17] = max_inv[:: 17]                    # some points outside their scale-invariance range
    pdesc = noisy(desc[src] if n else rng.integers(0, 256, (m, 32), dtype=np.uint8), rng, 0.06)
    normal = (ow[None, :].astype(np.float64) - Xw) if False else (Xw.astype(np.float64) - ow.astype(np.float64))
    normal /= np.maximum(np.linalg.norm(normal, axis=1, keepdims=True), 1e-9)
    normal += rng.normal(0, 0.5, normal.shape)             # some fail the 60 degree viewing test
    normal /= np.maximum(np.linalg.norm(normal, axis=1, keepdims=True), 1e-9)
    pv = N.make_posepoint_view(rng.random(m) < 0.93, Xw, pdesc, min_inv, max_inv, max_raw, rng.uniform(0, 360, m).astype(np.float32),
                               normal.astype(np.float32), keep)
    return fv, pv, ow, float(log_sf), keep


def bow_case(rng, nA, nB, n_nodes, N, mode):
    """Two feature sets with FeatureVectors over a shared vocabulary of n_nodes nodes."""
    keep = []
    descB = rng.integers(0, 256, (nB, 32), dtype=np.uint8)
    src = rng.integers(0, max(nB, 1), nA)
    descA = noisy(descB[src], rng, 0.07) if nB else rng.integers(0, 256, (nA, 32), dtype=np.uint8)
    nodeB = rng.integers(0, n_nodes, nB) * 7 + 3            # sparse, unordered node ids
    nodeA = np.where(rng.random(nA) < 0.85, nodeB[src] if nB else 0, rng.integers(0, n_nodes, nA) * 7 + 3)
    fvA, fvB = {}, {}
    for i in rng.permutation(nA) if nA else []:
        fvA.setdefault(int(nodeA[i]), []).append(int(i))
    for i in rng.permutation(nB) if nB else []:
        fvB.setdefault(int(nodeB[i]), []).append(int(i))
    angB = rng.uniform(0, 360, nB).astype(np.float32)
    angA = ((angB[src] if nB else np.zeros(nA)) + np.where(rng.random(nA) < 0.8, 20.0, rng.uniform(0, 360, nA))).astype(np.float32) % np.float32(360)
    a = N.make_bow_view(angA, descA, rng.random(nA) < 0.9, fvA, keep)
    b = N.make_bow_view(angB, descB, (rng.random(nB) < 0.9) if mode == 1 else None, fvB, keep)
    return a, b, keep


# ---------------------------------------------------------------------------------------------------------------
# E rows: Fuse, SearchBySim3, SearchForInitialization, SearchForTriangulation, ComputeDistinctiveDescriptors
# ---------------------------------------------------------------------------------------------------------------
def _pose(rng, rot_sigma=0.03, t_sigma=0.05):
    T = np.eye(4, dtype=np.float64)
    a = rng.normal(0, rot_sigma, 3)
    Rx = np.array([[1, 0, 0], [0, np.cos(a[0]), -np.sin(a[0])], [0, np.sin(a[0]), np.cos(a[0])]])
    Ry = np.array([[np.cos(a[1]), 0, np.sin(a[1])], [0, 1, 0], [-np.sin(a[1]), 0, np.cos(a[1])]])
    Rz = np.array([[np.cos(a[2]), -np.sin(a[2]), 0], [np.sin(a[2]), np.cos(a[2]), 0], [0, 0, 1]])
    T[:3, :3] = Rx @ Ry @ Rz
    T[:3, 3] = rng.normal(0, t_sigma, 3)
    return T.astype(np.float32)


def _project(T, Xw, K):
    Xc = Xw.astype(np.float64) @ T[:3, :3].astype(np.float64).T + T[:3, 3].astype(np.float64)
    return K["fx"] * Xc[:, 0] / Xc[:, 2] + K["cx"], K["fy"] * Xc[:, 1] / Xc[:, 2] + K["cy"], Xc[:, 2]


def _keyframe_of_points(rng, T, Xw, lvl, base_desc, K, N, extra):
    """Features = noisy projections of the world points (feature i <-> point i) + `extra` random ones."""
    m = len(Xw)
    u, v, z = _project(T, Xw, K)
    n = m + extra
    kp = np.zeros(n, N.KP_DTYPE)
    kp["x"][:m] = (u + rng.normal(0, 1.0, m)).astype(np.float32)
    kp["y"][:m] = (v + rng.normal(0, 1.0, m)).astype(np.float32)
    kp["x"][m:] = rng.uniform(0, 640, extra).astype(np.float32)
    kp["y"][m:] = rng.uniform(0, 480, extra).astype(np.float32)
    kp["octave"][:m] = np.clip(lvl - (rng.random(m) < 0.3), 0, 7)
    kp["octave"][m:] = rng.integers(0, 8, extra)
    kp["angle"] = rng.uniform(0, 360, n).astype(np.float32)
    desc = np.concatenate([noisy(base_desc, rng, 0.05), rng.integers(0, 256, (extra, 32), dtype=np.uint8)])
    ur = np.where(rng.random(n) < 0.7, kp["x"] - np.concatenate([K["bf"] / np.maximum(z, 0.1), rng.uniform(5, 40, extra)]), -1).astype(np.float32)
    return kp, desc, ur


def _points_view(rng, N, Xw, desc, lvl, cam_centre, valid_frac, keep, n_total=None):
    """pl_posepoint_view of map points seen from cam_centre; padded with invalid rows up to n_total."""
    m = len(Xw)
    n_total = m if n_total is None else n_total
    dist = np.linalg.norm(Xw.astype(np.float64) - cam_centre.astype(np.float64), axis=1)
    max_raw = (dist * np.float32(1.2) ** (lvl + rng.uniform(-0.45, 0.0, m))).astype(np.float32)   # PredictScale -> lvl
    max_inv = (np.float32(1.2) * max_raw).astype(np.float32)
    min_inv = (np.float32(0.8) * max_raw / np.float32(1.2) ** 7).astype(np.float32)
    if m:
        min_inv[:: 23] = max_inv[:: 23]
    normal = Xw.astype(np.float64) - cam_centre.astype(np.float64)
    normal /= np.maximum(np.linalg.norm(normal, axis=1, keepdims=True), 1e-9)
    normal += rng.normal(0, 0.35, normal.shape)
    normal /= np.maximum(np.linalg.norm(normal, axis=1, keepdims=True), 1e-9)
    pad = n_total - m

    def padded(a, fill=0):
        return np.concatenate([a, np.full((pad,) + a.shape[1:], fill, a.dtype)]) if pad else a
    valid = padded((rng.random(m) < valid_frac).astype(np.uint8))
    return N.make_posepoint_view(valid, padded(Xw.astype(np.float32)), padded(noisy(desc, rng, 0.04)), padded(min_inv), padded(max_inv, 1),
                                 padded(max_raw, 1), None, padded(normal.astype(np.float32)), keep)


def _centre(T):
    """Ow = -R^T t as the cv::Mat expression evaluates it (double accumulation, rounded once)."""
    return np.array([np.float32(-(sum(np.float64(T[k, r]) * np.float64(T[k, 3]) for k in range(3)))) for r in range(3)], np.float32)


def _world_points(rng, m, K):
    z = rng.uniform(0.8, 7, m)
    if m:
        z[: max(m // 50, 1)] *= -1
    u, v = rng.uniform(-40, 680, m), rng.uniform(-40, 520, m)
    return np.stack([(u - K["cx"]) * z / K["fx"], (v - K["cy"]) * z / K["fy"], z], 1).astype(np.float32)


def fuse_case(rng, m, extra, N, K, sf):
    """A key frame whose first m features are projections of m map points.  Returns (kf_view, pt_view, ow, log_sf, inv_sigma2, keep)."""
    keep = []
    T = _pose(rng)
    Xw = _world_points(rng, m, K)
    lvl = rng.integers(0, 8, m)
    base = rng.integers(0, 256, (m, 32), dtype=np.uint8)
    kp, desc, ur = _keyframe_of_points(rng, T, Xw, lvl, base, K, N, extra)
    fv = N.make_frame_view(kp, desc, ur, None, (0, 0, 640, 480), K, T[:3].reshape(-1), sf, keep)
    ow = _centre(T)
    pv = _points_view(rng, N, Xw, base, lvl, ow, 0.9, keep)
    inv_sigma2 = (1.0 / (np.asarray(sf, np.float32) ** 2)).astype(np.float32)
    return fv, pv, ow, float(np.float32(np.log(np.float32(1.2)))), inv_sigma2, keep


def sim3_case(rng, m, extra1, extra2, N, K, sf):
    """Two key frames observing the same m points.  Returns (kf1, kf2, pts1, pts2, t21, t12, log_sf, keep)."""
    keep = []
    T1, T2 = _pose(rng), _pose(rng)
    Xw = _world_points(rng, m, K)
    lvl = rng.integers(0, 8, m)
    base = rng.integers(0, 256, (m, 32), dtype=np.uint8)
    perm = rng.permutation(m)
    kp1, d1, ur1 = _keyframe_of_points(rng, T1, Xw, lvl, base, K, N, extra1)
    kp2, d2, ur2 = _keyframe_of_points(rng, T2, Xw[perm], lvl[perm], base[perm], K, N, extra2)
    kf1 = N.make_frame_view(kp1, d1, ur1, None, (0, 0, 640, 480), K, T1[:3].reshape(-1), sf, keep)
    kf2 = N.make_frame_view(kp2, d2, ur2, None, (0, 0, 640, 480), K, T2[:3].reshape(-1), sf, keep)
    # map points in camera distance terms: SearchBySim3 measures |p3Dc| in the OTHER camera
    pts1 = _points_view(rng, N, Xw, base, lvl, _centre(T2), 0.85, keep, n_total=m + extra1)
    pts2 = _points_view(rng, N, Xw[perm], base[perm], lvl[perm], _centre(T1), 0.85, keep, n_total=m + extra2)
    # p1 = R12 p2 + t12 with T12 = T1 * inv(T2), s12 = 1 (+ a small perturbation so the projections are not exact)
    T12 = T1.astype(np.float64) @ np.linalg.inv(T2.astype(np.float64))
    T12[:3, 3] += rng.normal(0, 0.002, 3)
    s12 = 1.02
    sR12 = s12 * T12[:3, :3]
    sR21 = (1.0 / s12) * T12[:3, :3].T
    t12 = T12[:3, 3]
    t21 = -sR21 @ t12
    t21m = np.concatenate([sR21, t21[:, None]], 1).astype(np.float32)
    t12m = np.concatenate([sR12, t12[:, None]], 1).astype(np.float32)
    return kf1, kf2, pts1, pts2, t21m, t12m, float(np.float32(np.log(np.float32(1.2)))), keep


def init_case(rng, n1, n2, N, K, sf):
    """Two monocular frames; F2 = F1 shifted by a few pixels + new features; most features on level 0."""
    keep = []
    kp1, d1, ur1 = rand_frame(rng, n1, N)
    kp1["octave"] = np.where(rng.random(n1) < 0.7, 0, rng.integers(1, 8, n1))
    src = rng.integers(0, max(n1, 1), n2)
    kp2 = np.zeros(n2, N.KP_DTYPE)
    if n1 and n2:
        kp2["x"] = (kp1["x"][src] + rng.normal(3, 6, n2)).astype(np.float32)
        kp2["y"] = (kp1["y"][src] + rng.normal(-2, 6, n2)).astype(np.float32)
        kp2["octave"] = np.where(rng.random(n2) < 0.9, kp1["octave"][src], rng.integers(0, 8, n2))
        kp2["angle"] = ((kp1["angle"][src] + np.where(rng.random(n2) < 0.8, 12.0, rng.uniform(0, 360, n2))) % 360).astype(np.float32)
        d2 = noisy(d1[src], rng, 0.06)
    else:
        kp2["x"] = rng.uniform(0, 640, n2).astype(np.float32)
        kp2["y"] = rng.uniform(0, 480, n2).astype(np.float32)
        d2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
    I = np.eye(4, dtype=np.float32)[:3].reshape(-1)
    f1 = N.make_frame_view(kp1, d1, np.full(n1, -1, np.float32), None, (0, 0, 640, 480), K, I, sf, keep)
    f2 = N.make_frame_view(kp2, d2, np.full(n2, -1, np.float32), None, (0, 0, 640, 480), K, I, sf, keep)
    prev = np.stack([kp1["x"], kp1["y"]], 1).astype(np.float32) if n1 else np.zeros((0, 2), np.float32)
    return f1, f2, prev, keep


def triang_case(rng, m, extra1, extra2, n_nodes, N, K, sf):
    """Two key frames with unmatched features of the same points, feature vectors, and the fundamental matrix F12 of their
    poses.  Returns (a, b, f12, cw1, tcw2, sigma2, keep)."""
    keep = []
    T1, T2 = _pose(rng, 0.02, 0.3), _pose(rng, 0.02, 0.3)
    Xw = _world_points(rng, m, K)
    lvl = rng.integers(0, 8, m)
    base = rng.integers(0, 256, (m, 32), dtype=np.uint8)
    perm = rng.permutation(m)
    kp1, d1, ur1 = _keyframe_of_points(rng, T1, Xw, lvl, base, K, N, extra1)
    kp2, d2, ur2 = _keyframe_of_points(rng, T2, Xw[perm], lvl[perm], base[perm], K, N, extra2)
    node_of_point = rng.integers(0, n_nodes, m) * 5 + 2
    node1 = np.concatenate([node_of_point, rng.integers(0, n_nodes, extra1) * 5 + 2])
    node2 = np.concatenate([np.where(rng.random(m) < 0.9, node_of_point[perm], rng.integers(0, n_nodes, m) * 5 + 2),
                            rng.integers(0, n_nodes, extra2) * 5 + 2])
    fv1, fv2 = {}, {}
    for i in rng.permutation(len(node1)):
        fv1.setdefault(int(node1[i]), []).append(int(i))
    for i in rng.permutation(len(node2)):
        fv2.setdefault(int(node2[i]), []).append(int(i))
    a = N.make_triang_view(kp1, d1, ur1, rng.random(len(kp1)) < 0.8, fv1, keep)
    b = N.make_triang_view(kp2, d2, ur2, rng.random(len(kp2)) < 0.8, fv2, keep)
    # F12 = K^-T [t12]x R12 K^-1 (LocalMapping::ComputeF12), evaluated in double and rounded
    T12 = T1.astype(np.float64) @ np.linalg.inv(T2.astype(np.float64))
    R12, t12 = T12[:3, :3], T12[:3, 3]
    tx = np.array([[0, -t12[2], t12[1]], [t12[2], 0, -t12[0]], [-t12[1], t12[0], 0]])
    Km = np.array([[K["fx"], 0, K["cx"]], [0, K["fy"], K["cy"]], [0, 0, 1]], np.float64)
    f12 = (np.linalg.inv(Km).T @ tx @ R12 @ np.linalg.inv(Km)).astype(np.float32)
    sigma2 = (np.asarray(sf, np.float32) ** 2).astype(np.float32)
    return a, b, f12, _centre(T1), T2[:3].reshape(-1), sigma2, keep


def distinctive_case(rng, sizes):
    """Groups of observations of the same feature (noisy copies of one descriptor, a few outliers)."""
    rows, off = [], [0]
    for n in sizes:
        base = rng.integers(0, 256, (1, 32), dtype=np.uint8)
        d = noisy(np.repeat(base, n, 0), rng, 0.08) if n else np.zeros((0, 32), np.uint8)
        if n > 3:
            d[rng.integers(0, n, max(n // 6, 1))] = rng.integers(0, 256, (max(n // 6, 1), 32), dtype=np.uint8)
        if n > 2:
            d[n - 1] = d[0]                                   # duplicated rows -> equal medians (first one must win)
        rows.append(d)
        off.append(off[-1] + n)
    return np.concatenate(rows) if rows else np.zeros((0, 32), np.uint8), np.asarray(off, np.int32)


# ---------------------------------------------------------------------------------------------------------------
# F rows: Frame glue
# ---------------------------------------------------------------------------------------------------------------
TUM1_DIST = np.array([0.262383, -0.953104, -0.005358, 0.002628, 1.163314], np.float32)   # Examples/RGB-D/TUM1.yaml:14-18


def frustum_case(rng, n_frames, m, K):
    """n_frames nearby poses and m map points around their common field of view."""
    T = np.stack([_pose(rng, 0.05, 0.1)[:3] for _ in range(n_frames)]) if n_frames else np.zeros((0, 3, 4), np.float32)
    ow = np.stack([_centre(np.vstack([t, [0, 0, 0, 1]]).astype(np.float32)) for t in T]) if n_frames else np.zeros((0, 3), np.float32)
    Xw = _world_points(rng, m, K)
    centre = ow.mean(0) if n_frames else np.zeros(3, np.float32)
    d = np.linalg.norm(Xw.astype(np.float64) - centre, axis=1)
    lvl = rng.integers(0, 8, m)
    max_raw = (d * np.float32(1.2) ** (lvl + rng.uniform(-0.5, 0.5, m))).astype(np.float32)
    max_inv = (np.float32(1.2) * max_raw).astype(np.float32)
    min_inv = (np.float32(0.8) * max_raw / np.float32(1.2) ** 7).astype(np.float32)
    if m:
        min_inv[:: 19] = max_inv[:: 19]
    normal = Xw.astype(np.float64) - centre
    normal /= np.maximum(np.linalg.norm(normal, axis=1, keepdims=True), 1e-9)
    normal += rng.normal(0, 0.6, normal.shape)
    normal /= np.maximum(np.linalg.norm(normal, axis=1, keepdims=True), 1e-9)
    return T.reshape(n_frames, 12).astype(np.float32), ow.astype(np.float32), Xw, normal.astype(np.float32), min_inv, max_inv, max_raw


# ---------------------------------------------------------------------------------------------------------------
# G row: a synthetic DBoW2 vocabulary (hierarchical clusters of binary descriptors) in loadFromTextFile's node order
# ---------------------------------------------------------------------------------------------------------------
def make_vocabulary(rng, k, L, leaf_depth_jitter=True, stop_frac=0.05):
    """Returns (parent, is_leaf, desc, weight) for nodes 1..n in breadth-first file order; children are noisy copies of the parent."""
    parent, leaf, desc, weight = [], [], [], []
    level = [(0, rng.integers(0, 256, 32, dtype=np.uint8))]
    for depth in range(1, L + 1):
        nxt = []
        for pid, pdesc in level:
            for _ in range(k):
                d = noisy(pdesc[None, :], rng, 0.18 / depth)[0]
                parent.append(pid)
                is_leaf = depth == L or (leaf_depth_jitter and depth >= 2 and rng.random() < 0.04)
                leaf.append(is_leaf)
                desc.append(d)
                weight.append(0.0 if (is_leaf and rng.random() < stop_frac) else float(rng.uniform(0.5, 9.0)) if is_leaf else 0.0)
                nid = len(parent)
                if not is_leaf:
                    nxt.append((nid, d))
        level = nxt
    return np.asarray(parent, np.int32), np.asarray(leaf, np.uint8), np.stack(desc), np.asarray(weight, np.float64)


def write_vocabulary_text(path, k, L, scoring, weighting, parent, leaf, desc, weight):
    """TemplatedVocabulary::saveToTextFile (TemplatedVocabulary.h:1426-1452) format."""
    with open(path, "w") as f:
        f.write(f"{k} {L}  {scoring} {weighting}\n")
        for i in range(len(parent)):
            f.write(f"{parent[i]} {int(leaf[i])} " + " ".join(str(int(b)) for b in desc[i]) + f" {float(weight[i])!r}\n")


def vocabulary_features(rng, desc, leaf, n):
    """n query descriptors: noisy copies of random leaves (so that several features share a word) + a few random rows."""
    leaves = np.nonzero(leaf)[0]
    src = leaves[rng.integers(0, len(leaves), n) % max(len(leaves) // 3, 1)]
    q = noisy(desc[src], rng, 0.03)
    if n > 4:
        q[:: 7] = rng.integers(0, 256, q[:: 7].shape, dtype=np.uint8)
    return q


def stereo_pair(synth, seed, w=640, h=480):
    """A rectified pair: the right image is the left one seen with a disparity that grows towards the bottom of the image
    (8 .. 40 px, piecewise constant in bands of rows, sub-pixel by linear interpolation) + independent sensor noise."""
    left = synth.frame(seed, w, h)
    rng = np.random.default_rng(seed)
    right = np.empty_like(left)
    xs = np.arange(w, dtype=np.float64)
    for y0 in range(0, h, 40):
        d = 8.0 + 32.0 * y0 / h + 0.37
        for y in range(y0, min(y0 + 40, h)):
            right[y] = np.clip(np.rint(np.interp(xs + d, xs, left[y].astype(np.float64))), 0, 255).astype(np.uint8)
    right = np.clip(right.astype(np.int16) + rng.integers(-2, 3, right.shape), 0, 255).astype(np.uint8)
    return left, right
