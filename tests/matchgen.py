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
