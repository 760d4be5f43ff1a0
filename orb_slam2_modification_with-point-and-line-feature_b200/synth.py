"""Deterministic synthetic inputs for the point+line front-end.

Pure numpy (no cv2) so the same frames can be produced on the GPU box.  All randomness is
``numpy.random.Generator(PCG64(seed))``.

* :func:`frame` — one textured grey image (configs 1, 3, 5): a grid of random-grey blocks with dark mortar lines,
  random filled rectangles and line strokes (to feed LSD), a small Gaussian blur and additive noise.
* :func:`room_sequence` — config 2: a three-plane "room" rendered with TUM1 intrinsics
  (reference Examples/RGB-D/TUM1.yaml:8-11) along a smooth camera trajectory, with float32 depth in metres.
* :func:`descriptor_sets` — config 4: train/query 32-byte descriptor sets with planted near-duplicates and ties.
"""
from __future__ import annotations

import numpy as np

TUM1 = dict(fx=517.306408, fy=516.469215, cx=318.643040, cy=255.313989, bf=40.0, th_depth=40.0)


def _blur_sep(img: np.ndarray, sigma: float) -> np.ndarray:
    r = max(1, int(np.ceil(3 * sigma)))
    k = np.exp(-0.5 * (np.arange(-r, r + 1) / sigma) ** 2)
    k /= k.sum()
    p = np.pad(img, ((0, 0), (r, r)), mode="reflect")
    out = np.zeros_like(img)
    for i, w in enumerate(k):
        out += w * p[:, i:i + img.shape[1]]
    p = np.pad(out, ((r, r), (0, 0)), mode="reflect")
    out2 = np.zeros_like(img)
    for i, w in enumerate(k):
        out2 += w * p[i:i + img.shape[0], :]
    return out2


def _stroke(img: np.ndarray, x0, y0, x1, y1, width, value):
    """Rasterise a thick line segment by distance-to-segment test inside its bounding box."""
    h, w = img.shape
    xa, xb = int(max(0, min(x0, x1) - width - 1)), int(min(w, max(x0, x1) + width + 2))
    ya, yb = int(max(0, min(y0, y1) - width - 1)), int(min(h, max(y0, y1) + width + 2))
    if xa >= xb or ya >= yb:
        return
    yy, xx = np.mgrid[ya:yb, xa:xb].astype(np.float64)
    dx, dy = x1 - x0, y1 - y0
    L2 = dx * dx + dy * dy + 1e-12
    t = np.clip(((xx - x0) * dx + (yy - y0) * dy) / L2, 0, 1)
    d2 = (xx - (x0 + t * dx)) ** 2 + (yy - (y0 + t * dy)) ** 2
    m = d2 <= (0.5 * width) ** 2 + 0.25
    img[ya:yb, xa:xb][m] = value


def texture(seed: int, width: int, height: int, grid=(40, 30), n_rect=60, n_strokes=40) -> np.ndarray:
    """Float64 texture in [0,255] before blur/noise."""
    rng = np.random.Generator(np.random.PCG64(seed))
    gx, gy = grid
    blocks = rng.uniform(30, 225, size=(gy, gx))
    ys = (np.arange(height) * gy // height)
    xs = (np.arange(width) * gx // width)
    img = blocks[np.ix_(ys, xs)].astype(np.float64)
    # 1-px dark mortar lines on block boundaries
    img[np.r_[False, ys[1:] != ys[:-1]], :] = 12.0
    img[:, np.r_[False, xs[1:] != xs[:-1]]] = 12.0
    for _ in range(n_rect):
        rw, rh = rng.integers(8, max(9, width // 6)), rng.integers(8, max(9, height // 6))
        x0, y0 = rng.integers(0, width - 4), rng.integers(0, height - 4)
        img[y0:y0 + rh, x0:x0 + rw] = rng.uniform(20, 235)
    for _ in range(n_strokes):
        length = rng.uniform(40, 300)
        ang = rng.uniform(0, np.pi)
        x0, y0 = rng.uniform(0, width), rng.uniform(0, height)
        x1, y1 = x0 + length * np.cos(ang), y0 + length * np.sin(ang)
        _stroke(img, x0, y0, x1, y1, float(rng.integers(1, 4)), rng.uniform(0, 255))
    return img


def frame(seed: int, width: int = 640, height: int = 480, grid=None) -> np.ndarray:
    """One synthetic uint8 grey frame (C-contiguous, shape (height, width))."""
    if grid is None:
        grid = (max(2, width // 16), max(2, height // 16))
    img = texture(seed, width, height, grid=grid)
    img = _blur_sep(img, 0.8)
    rng = np.random.Generator(np.random.PCG64(seed ^ 0x5EED))
    img = img + rng.normal(0.0, 2.0, size=img.shape)
    return np.ascontiguousarray(np.clip(np.rint(img), 0, 255).astype(np.uint8))


def frames(seed0: int, n: int, width: int = 640, height: int = 480) -> np.ndarray:
    """n frames: every 8th is rendered from scratch; the others are cheap deterministic variants (shift + noise)
    of the nearest rendered one — full rendering of thousands of frames would dominate bench start-up."""
    out = np.empty((n, height, width), np.uint8)
    base = None
    for i in range(n):
        if i % 8 == 0:
            base = frame(seed0 + i, width, height)
            out[i] = base
        else:
            rng = np.random.Generator(np.random.PCG64(seed0 + i))
            sx, sy = int(rng.integers(-24, 25)), int(rng.integers(-24, 25))
            v = np.roll(base, (sy, sx), axis=(0, 1)).astype(np.int16)
            v += rng.integers(-3, 4, size=v.shape, dtype=np.int16)
            out[i] = np.clip(v, 0, 255).astype(np.uint8)
    return out


# ---------------------------------------------------------------------------------------------------------------
# config 2: three-plane room, RGB-D sequence
# ---------------------------------------------------------------------------------------------------------------
def _rot(rx, ry, rz):
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def _frames_block_job(args):
    seed0, blk, lo, hi, width, height = args
    base = frame(seed0 + blk * 8, width, height)
    out = []
    for i in range(max(lo, blk * 8), min(hi, blk * 8 + 8)):
        if i % 8 == 0:
            out.append(base)
        else:
            rng = np.random.Generator(np.random.PCG64(seed0 + i))
            sx, sy = int(rng.integers(-24, 25)), int(rng.integers(-24, 25))
            v = np.roll(base, (sy, sx), axis=(0, 1)).astype(np.int16)
            v += rng.integers(-3, 4, size=v.shape, dtype=np.int16)
            out.append(np.clip(v, 0, 255).astype(np.uint8))
    return np.stack(out)


def frames_range(seed0: int, begin: int, end: int, width: int = 640, height: int = 480, workers: int = 1) -> np.ndarray:
    """frames(seed0, N, ...)[begin:end] for any N >= end, without rendering the rest: frame i only depends on seed0 and i (its block of
    eight), which is what lets every rank of a sharded run render its own contiguous range.  `workers` processes render the blocks."""
    if end <= begin:
        return np.empty((0, height, width), np.uint8)
    jobs = [(seed0, blk, begin, end, width, height) for blk in range(begin // 8, (end + 7) // 8)]
    if workers <= 1 or len(jobs) < 4:
        res = [_frames_block_job(j) for j in jobs]
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(min(workers, len(jobs))) as pool:
            res = pool.map(_frames_block_job, jobs, chunksize=max(1, len(jobs) // (4 * workers)))
    return np.concatenate(res)


def room_trajectory(n: int, seed: int = 2003) -> np.ndarray:
    """n world->camera poses Tcw (n,4,4) float64: smooth motion, <= 2 cm and <= 0.5 deg per frame."""
    rng = np.random.Generator(np.random.PCG64(seed))
    t = np.linspace(0, 1, n)
    ph = rng.uniform(0, 2 * np.pi, size=6)
    pos = np.stack([0.9 * np.sin(2 * np.pi * 0.35 * t + ph[0]) * 0.6,
                    0.15 * np.sin(2 * np.pi * 0.5 * t + ph[1]),
                    0.5 * np.sin(2 * np.pi * 0.3 * t + ph[2]) * 0.6], axis=1)
    ang = np.stack([np.deg2rad(4) * np.sin(2 * np.pi * 0.4 * t + ph[3]),
                    np.deg2rad(14) * np.sin(2 * np.pi * 0.3 * t + ph[4]),
                    np.deg2rad(3) * np.sin(2 * np.pi * 0.5 * t + ph[5])], axis=1)
    T = np.zeros((n, 4, 4))
    for i in range(n):
        Rwc = _rot(*ang[i])
        Rcw = Rwc.T
        T[i, :3, :3] = Rcw
        T[i, :3, 3] = -Rcw @ pos[i]
        T[i, 3, 3] = 1
    return T


class Room:
    """Front wall z=3 (4 m x 3 m visible region tiles further), left wall x=-2, floor y=1.5 (camera y down)."""

    def __init__(self, tex_px_per_m: int = 320):
        self.ppm = tex_px_per_m
        self.tex = [frame(2000 + i, 8 * tex_px_per_m // 2, 6 * tex_px_per_m // 2).astype(np.float32) for i in range(3)]
        # planes: (point, normal, u axis, v axis)
        self.planes = [
            (np.array([0, 0, 3.0]), np.array([0, 0, -1.0]), np.array([1.0, 0, 0]), np.array([0, 1.0, 0])),
            (np.array([-2.0, 0, 0]), np.array([1.0, 0, 0]), np.array([0, 0, 1.0]), np.array([0, 1.0, 0])),
            (np.array([0, 1.5, 0]), np.array([0, -1.0, 0]), np.array([1.0, 0, 0]), np.array([0, 0, 1.0])),
        ]

    def render(self, Tcw: np.ndarray, width=640, height=480, K=TUM1):
        Rcw, tcw = Tcw[:3, :3], Tcw[:3, 3]
        Rwc = Rcw.T
        Ow = -Rwc @ tcw
        u, v = np.meshgrid(np.arange(width, dtype=np.float64), np.arange(height, dtype=np.float64))
        dc = np.stack([(u - K["cx"]) / K["fx"], (v - K["cy"]) / K["fy"], np.ones_like(u)], axis=-1)
        dw = dc @ Rwc.T
        best_t = np.full((height, width), np.inf)
        img = np.zeros((height, width), np.float32)
        for (p0, nrm, ua, va), tex in zip(self.planes, self.tex):
            denom = dw @ nrm
            with np.errstate(divide="ignore", invalid="ignore"):
                t = ((p0 - Ow) @ nrm) / denom
            ok = (denom < -1e-9) & (t > 0.05) & (t < best_t)
            P = Ow + dw * t[..., None]
            tu = ((P - p0) @ ua) * self.ppm / 2 + tex.shape[1] / 2
            tv = ((P - p0) @ va) * self.ppm / 2 + tex.shape[0] / 2
            tu = np.mod(tu, tex.shape[1] - 1.001)
            tv = np.mod(tv, tex.shape[0] - 1.001)
            x0 = np.floor(tu).astype(np.int64)
            y0 = np.floor(tv).astype(np.int64)
            fx, fy = (tu - x0).astype(np.float32), (tv - y0).astype(np.float32)
            x0 = np.clip(x0, 0, tex.shape[1] - 2)
            y0 = np.clip(y0, 0, tex.shape[0] - 2)
            val = (tex[y0, x0] * (1 - fx) * (1 - fy) + tex[y0, x0 + 1] * fx * (1 - fy) +
                   tex[y0 + 1, x0] * (1 - fx) * fy + tex[y0 + 1, x0 + 1] * fx * fy)
            img = np.where(ok, val, img)
            best_t = np.where(ok, t, best_t)
        depth = np.where(np.isfinite(best_t) & (best_t < 8.0), best_t, 0.0).astype(np.float32)  # z along the ray dir
        gray = np.clip(np.rint(img), 0, 255).astype(np.uint8)
        return np.ascontiguousarray(gray), np.ascontiguousarray(depth)


_ROOM = None


def _render_job(args):
    global _ROOM
    if _ROOM is None:
        _ROOM = Room()
    T, width, height = args
    return _ROOM.render(T, width, height)


def room_sequence(n: int = 300, width: int = 640, height: int = 480, seed: int = 2003, workers: int = 1):
    """Returns (gray (n,h,w) uint8, depth (n,h,w) float32 metres, Tcw (n,4,4) float32).  `workers` > 1 renders the
    frames in a process pool (rendering is pure numpy and deterministic, so the result does not depend on it)."""
    T = room_trajectory(n, seed)
    gray = np.empty((n, height, width), np.uint8)
    depth = np.empty((n, height, width), np.float32)
    jobs = [(T[i], width, height) for i in range(n)]
    if workers > 1 and n > 4:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(min(workers, n)) as pool:
            res = pool.map(_render_job, jobs, chunksize=max(1, n // (4 * workers)))
    else:
        res = [_render_job(j) for j in jobs]
    for i, (g, d) in enumerate(res):
        gray[i], depth[i] = g, d
    return gray, depth, T.astype(np.float32)


# ---------------------------------------------------------------------------------------------------------------
# config 4: descriptor sets
# ---------------------------------------------------------------------------------------------------------------
def descriptor_sets(n: int, flip_p: float = 0.08, dup_block: int = 16):
    """(query, train): uint8 (n,32).  80 % of queries are noisy copies of a permuted train row, 20 % are fresh
    random; the first `dup_block` train rows are duplicated further down to exercise tie-breaks."""
    rt = np.random.Generator(np.random.PCG64(4000 + n))
    rq = np.random.Generator(np.random.PCG64(5000 + n))
    train = rt.integers(0, 256, size=(n, 32), dtype=np.uint8)
    if n >= 4 * dup_block:
        train[n // 2:n // 2 + dup_block] = train[:dup_block]
    perm = rq.permutation(n)
    q = train[perm].copy()
    flips = (rq.random(size=(n, 256)) < flip_p)
    q ^= np.packbits(flips, axis=1, bitorder="little")
    fresh = rq.random(n) < 0.2
    q[fresh] = rq.integers(0, 256, size=(int(fresh.sum()), 32), dtype=np.uint8)
    return np.ascontiguousarray(q), np.ascontiguousarray(train)
