"""ctypes binding of the CPU oracle (oracle/_build/libploracle.so).  TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs — never by
the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libploracle.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
KL_DTYPE = np.dtype([("angle", "<f4"), ("class_id", "<i4"), ("octave", "<i4"), ("pt_x", "<f4"), ("pt_y", "<f4"),
                     ("response", "<f4"), ("size", "<f4"), ("sx", "<f4"), ("sy", "<f4"), ("ex", "<f4"), ("ey", "<f4"),
                     ("sx_oct", "<f4"), ("sy_oct", "<f4"), ("ex_oct", "<f4"), ("ey_oct", "<f4"), ("length", "<f4"),
                     ("num_pixels", "<i4")])
assert KP_DTYPE.itemsize == 28 and KL_DTYPE.itemsize == 68


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h"))]
    srcs.append(os.path.join(_HERE, "..", "include", "plslam_c.h"))
    if force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_fast_atan2.restype = C.c_float
        _lib.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        _lib.orc_orb_create.restype = C.c_void_p
        _lib.orc_orb_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        _lib.orc_orb_destroy.argtypes = [C.c_void_p]
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _u8(img):
    img = np.asarray(img)
    assert img.dtype == np.uint8 and img.ndim == 2 and (img.size == 0 or img.strides[1] == 1)
    return img


def fast_atan2(y, x):
    y = np.asarray(y, np.float32)
    x = np.asarray(x, np.float32)
    f = lib().orc_fast_atan2
    return np.array([f(float(a), float(b)) for a, b in zip(y.ravel(), x.ravel())], np.float32).reshape(y.shape)


def resize_linear(img, dw, dh):
    img = _u8(img)
    out = np.empty((dh, dw), np.uint8)
    lib().orc_resize_linear_u8(_p(img), C.c_int(img.shape[1]), C.c_int(img.shape[0]), C.c_size_t(img.strides[0]), _p(out),
                               C.c_int(dw), C.c_int(dh), C.c_size_t(dw))
    return out


def border_reflect101(img, border):
    img = _u8(img)
    h, w = img.shape
    out = np.empty((h + 2 * border, w + 2 * border), np.uint8)
    lib().orc_border_reflect101_u8(_p(img), C.c_int(w), C.c_int(h), C.c_size_t(img.strides[0]), _p(out),
                                   C.c_size_t(out.strides[0]), C.c_int(border))
    return out


def gaussian_blur7(img):
    img = _u8(img)
    h, w = img.shape
    out = np.empty((h, w), np.uint8)
    lib().orc_gaussian_blur7_u8(_p(img), C.c_int(w), C.c_int(h), C.c_size_t(img.strides[0]), _p(out), C.c_size_t(w))
    return out


def gaussian_blur_fixed(img, kernel):
    img = _u8(img)
    h, w = img.shape
    out = np.empty((h, w), np.uint8)
    k = np.asarray(kernel, np.int32)
    lib().orc_gaussian_blur_fixed_u8(_p(img), C.c_int(w), C.c_int(h), C.c_size_t(img.strides[0]), _p(out), C.c_size_t(w),
                                     _p(k), C.c_int(len(k)))
    return out


def fast_detect(img, threshold, nonmax=True):
    img = _u8(img)
    h, w = img.shape
    cap = max(16, h * w)
    xs = np.empty(cap, np.float32)
    ys = np.empty(cap, np.float32)
    rs = np.empty(cap, np.float32)
    n = lib().orc_fast_detect(_p(img), C.c_int(w), C.c_int(h), C.c_size_t(img.strides[0]), C.c_int(threshold),
                              C.c_int(1 if nonmax else 0), _p(xs), _p(ys), _p(rs), C.c_int(cap))
    return xs[:n].copy(), ys[:n].copy(), rs[:n].copy()


def distribute_octtree(xs, ys, resp, minX, maxX, minY, maxY, N):
    xs = np.ascontiguousarray(xs, np.float32)
    ys = np.ascontiguousarray(ys, np.float32)
    resp = np.ascontiguousarray(resp, np.float32)
    n = len(xs)
    cap = n + 8
    ox = np.empty(cap, np.float32)
    oy = np.empty(cap, np.float32)
    orr = np.empty(cap, np.float32)
    m = lib().orc_distribute_octtree(_p(xs), _p(ys), _p(resp), C.c_int(n), C.c_int(minX), C.c_int(maxX), C.c_int(minY),
                                     C.c_int(maxY), C.c_int(N), _p(ox), _p(oy), _p(orr), C.c_int(cap))
    return ox[:m].copy(), oy[:m].copy(), orr[:m].copy()


class OrbOracle:
    """Mirror of ORB_SLAM2::ORBextractor on the CPU oracle."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.nlevels = nlevels
        self.nfeatures = nfeatures
        self._h = C.c_void_p(lib().orc_orb_create(nfeatures, C.c_float(scale_factor), nlevels, ini_th, min_th))

    def __del__(self):
        try:
            if self._h:
                lib().orc_orb_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def tables(self):
        n = self.nlevels
        sf, isf, s2, is2 = (np.empty(n, np.float32) for _ in range(4))
        per = np.empty(n, np.int32)
        umax = np.empty(16, np.int32)
        lib().orc_orb_tables(self._h, _p(sf), _p(isf), _p(s2), _p(is2), _p(per), _p(umax))
        return dict(scale_factors=sf, inv_scale_factors=isf, sigma2=s2, inv_sigma2=is2, per_level=per, umax=umax)

    def extract(self, img):
        img = _u8(img)
        cap = self.nfeatures + 4 * self.nlevels + 64
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        rc = lib().orc_orb_extract(self._h, _p(img), C.c_int(img.shape[0]), C.c_int(img.shape[1]),
                                   C.c_size_t(img.strides[0]), _p(kps), _p(desc), C.c_int(cap), C.byref(n))
        if rc == -2:
            return kps[:0], desc[:0]
        assert rc == 0, rc
        return kps[:n.value].copy(), desc[:n.value].copy()

    def level_dims(self, level):
        w, h = C.c_int(), C.c_int()
        assert lib().orc_orb_level_dims(self._h, C.c_int(level), C.byref(w), C.byref(h)) == 0
        return w.value, h.value

    def level_bordered(self, level):
        w, h = self.level_dims(level)
        out = np.empty((h + 38, w + 38), np.uint8)
        assert lib().orc_orb_level_bordered(self._h, C.c_int(level), _p(out)) == 0
        return out

    def level_blurred(self, level):
        w, h = self.level_dims(level)
        out = np.empty((h, w), np.uint8)
        rc = lib().orc_orb_level_blurred(self._h, C.c_int(level), _p(out))
        return out if rc == 0 else None

    def level_candidates(self, level):
        cap = 1 << 18
        xs = np.empty(cap, np.float32)
        ys = np.empty(cap, np.float32)
        rs = np.empty(cap, np.float32)
        n = lib().orc_orb_level_candidates(self._h, C.c_int(level), _p(xs), _p(ys), _p(rs), C.c_int(cap))
        return xs[:n].copy(), ys[:n].copy(), rs[:n].copy()


# ---------------------------------------------------------------------------------------------------------------
# descriptor search
# ---------------------------------------------------------------------------------------------------------------
def _rows(a):
    a = np.ascontiguousarray(a, np.uint8)
    assert a.ndim == 2 and a.shape[1] == 32
    return a


def hamming_pairs(a, b):
    a, b = _rows(a), _rows(b)
    out = np.empty(a.shape[0], np.int32)
    lib().orc_hamming_pairs(_p(a), _p(b), C.c_int(a.shape[0]), _p(out))
    return out


def hamming_knn2(q, t, threads=1):
    q, t = _rows(q), _rows(t)
    idx = np.empty((q.shape[0], 2), np.int32)
    dist = np.empty((q.shape[0], 2), np.int32)
    lib().orc_hamming_knn2(_p(q), C.c_int(q.shape[0]), _p(t), C.c_int(t.shape[0]), _p(idx), _p(dist), C.c_int(threads))
    return idx, dist


def hamming_candidates(q, t, off, cidx):
    q, t = _rows(q), _rows(t)
    off = np.ascontiguousarray(off, np.int32)
    cidx = np.ascontiguousarray(cidx, np.int32)
    out = np.empty(cidx.shape[0], np.int32)
    lib().orc_hamming_candidates(_p(q), C.c_int(q.shape[0]), _p(t), _p(off), _p(cidx), _p(out))
    return out


# ---------------------------------------------------------------------------------------------------------------
# line extraction
# ---------------------------------------------------------------------------------------------------------------
def lsd_detect(img, order_mode=0, cap=20000):
    """cv::createLineSegmentDetector(LSD_REFINE_ADV).detect -> (lines (n,4) f32, width, prec, nfa f64)."""
    img = _u8(img)
    xy = np.empty((cap, 4), np.float32)
    w, p, nf = (np.empty(cap, np.float64) for _ in range(3))
    n = lib().orc_lsd_detect(_p(img), C.c_int(img.shape[0]), C.c_int(img.shape[1]), C.c_size_t(img.strides[0]),
                             C.c_int(order_mode), _p(xy), _p(w), _p(p), _p(nf), C.c_int(cap))
    assert n <= cap
    return xy[:n].copy(), w[:n].copy(), p[:n].copy(), nf[:n].copy()


def lsd_trace(img, cap=60000):
    """The regions flsd() grows, in order: rows of (seed pixel of the scaled image, first growth size, final size, got a rectangle)."""
    img = _u8(img)
    out = np.empty((cap, 4), np.int32)
    n = lib().orc_lsd_trace(_p(img), C.c_int(img.shape[0]), C.c_int(img.shape[1]), C.c_size_t(img.strides[0]), _p(out), C.c_int(cap))
    assert n <= cap
    return out[:n].copy()


def lsd_scaled(img):
    img = _u8(img)
    ow, oh = C.c_int(), C.c_int()
    lib().orc_lsd_scaled(_p(img), C.c_int(img.shape[0]), C.c_int(img.shape[1]), C.c_size_t(img.strides[0]), None,
                         C.byref(ow), C.byref(oh))
    out = np.empty((oh.value, ow.value), np.uint8)
    lib().orc_lsd_scaled(_p(img), C.c_int(img.shape[0]), C.c_int(img.shape[1]), C.c_size_t(img.strides[0]), _p(out),
                         C.byref(ow), C.byref(oh))
    return out


def lbd_compute(img, kls):
    img = _u8(img)
    kls = np.ascontiguousarray(kls, KL_DTYPE)
    n = len(kls)
    desc = np.zeros((n, 32), np.uint8)
    fdesc = np.zeros((n, 72), np.float32)
    lib().orc_lbd_compute(_p(img), C.c_int(img.shape[0]), C.c_int(img.shape[1]), C.c_size_t(img.strides[0]), _p(kls),
                          C.c_int(n), _p(desc), _p(fdesc))
    return desc, fdesc


def line_extract(img, max_lines=80, order_mode=0):
    """LineExtractor::ExtractLineSegment -> (keylines[KL_DTYPE], desc (n,32), coeffs (n,3) f64)."""
    img = _u8(img)
    cap = max(max_lines, 1)
    kls = np.zeros(cap, KL_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    co = np.zeros((cap, 3), np.float64)
    n = C.c_int(0)
    rc = lib().orc_line_extract(_p(img), C.c_int(img.shape[0]), C.c_int(img.shape[1]), C.c_size_t(img.strides[0]),
                                C.c_int(max_lines), C.c_int(order_mode), _p(kls), _p(desc), _p(co), C.byref(n))
    if rc == -2:
        return kls[:0], desc[:0], co[:0]
    assert rc == 0, rc
    return kls[:n.value].copy(), desc[:n.value].copy(), co[:n.value].copy()


# ---------------------------------------------------------------------------------------------------------------
# projection searches / line matching (views are the ctypes structures of the product's _native module)
# ---------------------------------------------------------------------------------------------------------------
def search_last_frame(cur_view, last_view, th, mono=False, check_orientation=True):
    match = np.empty(max(cur_view.n, 1), np.int32)
    n = C.c_int(0)
    lib().orc_orb_search_last_frame(C.byref(cur_view), C.byref(last_view), C.c_float(th), C.c_int(int(mono)),
                                    C.c_int(int(check_orientation)), _p(match), C.byref(n))
    return match[:cur_view.n], n.value


def search_local_points(frame_view, mp_view, th, nn_ratio):
    match = np.empty(max(frame_view.n, 1), np.int32)
    n = C.c_int(0)
    lib().orc_orb_search_local_points(C.byref(frame_view), C.byref(mp_view), C.c_float(th), C.c_float(nn_ratio), _p(match), C.byref(n))
    return match[:frame_view.n], n.value


def project_lines(start3d, end3d, src_kl, valid, tcw, K, bounds, img_size):
    s3 = np.ascontiguousarray(start3d, np.float64).reshape(-1, 3)
    e3 = np.ascontiguousarray(end3d, np.float64).reshape(-1, 3)
    kl = np.ascontiguousarray(src_kl, KL_DTYPE)
    va = np.ascontiguousarray(valid, np.uint8)
    n = len(kl)
    out = np.zeros(max(n, 1), KL_DTYPE)
    idx = np.zeros(max(n, 1), np.int32)
    m = C.c_int(0)
    t = (C.c_float * 12)(*[float(x) for x in np.asarray(tcw, np.float32).reshape(-1)[:12]])
    lib().orc_line_project(_p(s3), _p(e3), _p(kl), _p(va), C.c_int(n), t, C.c_float(K["fx"]), C.c_float(K["fy"]), C.c_float(K["cx"]),
                           C.c_float(K["cy"]), C.c_float(bounds[0]), C.c_float(bounds[1]), C.c_float(bounds[2]), C.c_float(bounds[3]),
                           C.c_int(img_size[0]), C.c_int(img_size[1]), _p(out), _p(idx), C.byref(m))
    return out[:m.value].copy(), idx[:m.value].copy()


def match_lines(proj_kl, proj_desc, cur_kl, cur_desc, cur_claimed=None):
    pk = np.ascontiguousarray(proj_kl, KL_DTYPE)
    pd = np.ascontiguousarray(proj_desc, np.uint8).reshape(-1, 32)
    ck = np.ascontiguousarray(cur_kl, KL_DTYPE)
    cd = np.ascontiguousarray(cur_desc, np.uint8).reshape(-1, 32)
    cc = None if cur_claimed is None else np.ascontiguousarray(cur_claimed, np.uint8)
    match = np.full(max(len(ck), 1), -1, np.int32)
    n, rel = C.c_int(0), C.c_int(0)
    lib().orc_line_match_pairs(_p(pk), _p(pd), C.c_int(len(pk)), _p(ck), _p(cd), None if cc is None else _p(cc), C.c_int(len(ck)),
                               _p(match), C.byref(n), C.byref(rel))
    return match[:len(ck)], n.value, rel.value


def line_iterator_count(x0, y0, x1, y1, cols, rows):
    lib().orc_line_iterator_count.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int]
    return lib().orc_line_iterator_count(x0, y0, x1, y1, cols, rows)


def line_search_by_projection(cur_view, line_view):
    """LineMatcher::SearchByProjection on the POD views: project + clip, then all-pairs LineMatching.
    Returns (match_of_line -> original map-line index or -1, nmatches, used_relaxed, n_projected)."""
    n = line_view.n
    out = np.zeros(max(n, 1), KL_DTYPE)
    idx = np.zeros(max(n, 1), np.int32)
    m = C.c_int(0)
    lib().orc_line_project(C.c_void_p(line_view.start3d), C.c_void_p(line_view.end3d), C.c_void_p(line_view.kl), C.c_void_p(line_view.valid),
                           C.c_int(n), cur_view.tcw, C.c_float(cur_view.fx), C.c_float(cur_view.fy), C.c_float(cur_view.cx), C.c_float(cur_view.cy),
                           C.c_float(cur_view.min_x), C.c_float(cur_view.min_y), C.c_float(cur_view.max_x), C.c_float(cur_view.max_y),
                           C.c_int(cur_view.cols), C.c_int(cur_view.rows), _p(out), _p(idx), C.byref(m))
    npj = m.value
    desc = np.ctypeslib.as_array((C.c_uint8 * (max(n, 1) * 32)).from_address(line_view.desc)).reshape(-1, 32) if n else np.zeros((0, 32), np.uint8)
    pdesc = np.ascontiguousarray(desc[idx[:npj]]) if npj else np.zeros((0, 32), np.uint8)
    match = np.full(max(cur_view.n, 1), -1, np.int32)
    cnt, rel = C.c_int(0), C.c_int(0)
    if cur_view.n:
        lib().orc_line_match_pairs(_p(out), _p(pdesc) if npj else None, C.c_int(npj), C.c_void_p(cur_view.kl), C.c_void_p(cur_view.desc),
                                   C.c_void_p(cur_view.claimed), C.c_int(cur_view.n), _p(match), C.byref(cnt), C.byref(rel))
    match = match[:cur_view.n]
    res = np.where(match >= 0, idx[np.clip(match, 0, max(npj - 1, 0))], -1).astype(np.int32) if npj else match.copy()
    return res, cnt.value, rel.value, npj


def search_keyframe_points(cur_view, pt_view, ow, log_sf, th, orb_dist, check_orientation=True):
    match = np.empty(max(cur_view.n, 1), np.int32)
    n = C.c_int(0)
    o = np.ascontiguousarray(ow, np.float32)
    lib().orc_orb_search_keyframe_points(C.byref(cur_view), C.byref(pt_view), _p(o), C.c_float(log_sf), C.c_float(th), C.c_int(int(orb_dist)),
                                         C.c_int(int(check_orientation)), _p(match), C.byref(n))
    return match[:cur_view.n], n.value


def search_sim3_points(kf_view, pt_view, ow, log_sf, th):
    match = np.empty(max(kf_view.n, 1), np.int32)
    n = C.c_int(0)
    o = np.ascontiguousarray(ow, np.float32)
    lib().orc_orb_search_sim3_points(C.byref(kf_view), C.byref(pt_view), _p(o), C.c_float(log_sf), C.c_int(int(th)), _p(match), C.byref(n))
    return match[:kf_view.n], n.value


def search_bow(a_view, b_view, mode, nn_ratio, check_orientation=True):
    sz = b_view.n if mode == 0 else a_view.n
    match = np.empty(max(sz, 1), np.int32)
    n = C.c_int(0)
    lib().orc_orb_search_bow(C.byref(a_view), C.byref(b_view), C.c_int(mode), C.c_float(nn_ratio), C.c_int(int(check_orientation)), _p(match),
                             C.byref(n))
    return match[:sz], n.value


def line_match_knn_ratio(ref_desc, cur_desc):
    r, c = _rows(ref_desc), _rows(cur_desc)
    match = np.full(max(c.shape[0], 1), -1, np.int32)
    n = C.c_int(0)
    lib().orc_line_match_knn_ratio(_p(r), C.c_int(r.shape[0]), _p(c), C.c_int(c.shape[0]), _p(match), C.byref(n))
    return match[:c.shape[0]], n.value


def line_search_for_triangulation(desc1, desc2):
    a, b = _rows(desc1), _rows(desc2)
    pairs = np.zeros((max(a.shape[0], 1), 2), np.int32)
    n = C.c_int(0)
    m1, m2 = C.c_double(0), C.c_double(0)
    lib().orc_line_search_for_triangulation(_p(a), C.c_int(a.shape[0]), _p(b), C.c_int(b.shape[0]), _p(pairs), C.byref(n), C.byref(m1), C.byref(m2))
    return pairs[:n.value].copy(), m1.value, m2.value


def line_fuse_candidates(ml_desc, valid, kf_desc):
    a, b = _rows(ml_desc), _rows(kf_desc)
    va = None if valid is None else np.ascontiguousarray(valid, np.uint8)
    tdx = np.full(max(a.shape[0], 1), -1, np.int32)
    n = C.c_int(0)
    lib().orc_line_fuse_candidates(_p(a), _p(va) if va is not None else None, C.c_int(a.shape[0]), _p(b), C.c_int(b.shape[0]), _p(tdx), C.byref(n))
    return tdx[:a.shape[0]], n.value


def fuse_candidates(kf_view, pt_view, ow, log_sf, inv_level_sigma2, th, variant=0):
    bi = np.empty(max(pt_view.n, 1), np.int32)
    bd = np.empty(max(pt_view.n, 1), np.int32)
    n = C.c_int(0)
    o = np.ascontiguousarray(ow, np.float32)
    sg = None if inv_level_sigma2 is None else np.ascontiguousarray(inv_level_sigma2, np.float32)
    lib().orc_orb_fuse_candidates(C.byref(kf_view), C.byref(pt_view), _p(o), C.c_float(log_sf), _p(sg) if sg is not None else None, C.c_float(th),
                                  C.c_int(variant), _p(bi), _p(bd), C.byref(n))
    return bi[:pt_view.n], bd[:pt_view.n], n.value


def search_by_sim3(kf1, kf2, pts1, pts2, t21, t12, log_sf1, log_sf2, th):
    t21 = np.ascontiguousarray(t21, np.float32).reshape(-1)[:12].copy()
    t12 = np.ascontiguousarray(t12, np.float32).reshape(-1)[:12].copy()
    m = np.empty(max(kf1.n, 1), np.int32)
    n = C.c_int(0)
    lib().orc_orb_search_by_sim3(C.byref(kf1), C.byref(kf2), C.byref(pts1), C.byref(pts2), _p(t21), _p(t12), C.c_float(log_sf1), C.c_float(log_sf2),
                                 C.c_float(th), _p(m), C.byref(n))
    return m[:kf1.n], n.value


def search_for_initialization(f1, f2, prev_matched, window_size, nn_ratio=0.9, check_orientation=True):
    pm = np.array(prev_matched, np.float32).reshape(-1, 2).copy()
    m = np.empty(max(f1.n, 1), np.int32)
    n = C.c_int(0)
    lib().orc_orb_search_for_initialization(C.byref(f1), C.byref(f2), _p(pm), C.c_int(int(window_size)), C.c_float(nn_ratio),
                                            C.c_int(int(check_orientation)), _p(m), C.byref(n))
    return m[:f1.n], n.value, pm


def search_for_triangulation(a, b, f12, cw1, kf2_tcw, K2, scale_factors2, level_sigma2_2, only_stereo=False, check_orientation=True):
    f12 = np.ascontiguousarray(f12, np.float32).reshape(-1)[:9].copy()
    cw = np.ascontiguousarray(cw1, np.float32).reshape(-1)[:3].copy()
    t2 = np.ascontiguousarray(kf2_tcw, np.float32).reshape(-1)[:12].copy()
    sf = np.ascontiguousarray(scale_factors2, np.float32)
    sg = np.ascontiguousarray(level_sigma2_2, np.float32)
    pairs = np.empty((max(a.bow.n, 1), 2), np.int32)
    n = C.c_int(0)
    lib().orc_orb_search_for_triangulation(C.byref(a), C.byref(b), _p(f12), _p(cw), _p(t2), C.c_float(K2["fx"]), C.c_float(K2["fy"]),
                                           C.c_float(K2["cx"]), C.c_float(K2["cy"]), _p(sf), _p(sg), C.c_int(len(sf)), C.c_int(int(only_stereo)),
                                           C.c_int(int(check_orientation)), _p(pairs), C.byref(n))
    return pairs[:max(n.value, 0)], n.value


def distinctive_descriptors(desc, group_off):
    d = _rows(desc)
    off = np.ascontiguousarray(group_off, np.int32)
    best = np.empty(max(len(off) - 1, 1), np.int32)
    lib().orc_distinctive_descriptors(_p(d), _p(off), C.c_int(len(off) - 1), _p(best))
    return best[:len(off) - 1]


def frame_undistort_points(xy, K, dist_coef):
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    dc = np.ascontiguousarray(dist_coef, np.float32).reshape(-1)[:5].copy()
    out = np.empty_like(xy)
    lib().orc_frame_undistort_points(_p(xy), C.c_int(len(xy)), C.c_float(K["fx"]), C.c_float(K["fy"]), C.c_float(K["cx"]), C.c_float(K["cy"]), _p(dc),
                                     _p(out))
    return out


def frame_stereo_from_rgbd_batch(depth, off, xy, x_un, bf):
    depth = np.ascontiguousarray(depth, np.float32)
    nf, rows, cols = depth.shape
    off = np.ascontiguousarray(off, np.int32)
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    xu = np.ascontiguousarray(x_un, np.float32)
    d = np.empty(max(len(xu), 1), np.float32)
    ur = np.empty(max(len(xu), 1), np.float32)
    lib().orc_frame_stereo_from_rgbd_batch(C.c_int(nf), _p(depth), C.c_int(rows), C.c_int(cols), C.c_size_t(cols * 4), C.c_size_t(rows * cols * 4),
                                           _p(off), _p(xy), _p(xu), C.c_float(bf), _p(d), _p(ur))
    return d[:len(xu)], ur[:len(xu)]


def frame_unproject_batch(off, xy_un, z, rwc, ow, K):
    off = np.ascontiguousarray(off, np.int32)
    xy = np.ascontiguousarray(xy_un, np.float32).reshape(-1, 2)
    z = np.ascontiguousarray(z, np.float32)
    rwc = np.ascontiguousarray(rwc, np.float32).reshape(-1, 9)
    ow = np.ascontiguousarray(ow, np.float32).reshape(-1, 3)
    w = np.empty((max(len(z), 1), 3), np.float32)
    v = np.empty(max(len(z), 1), np.uint8)
    lib().orc_frame_unproject_batch(C.c_int(len(off) - 1), _p(off), _p(xy), _p(z), _p(rwc), _p(ow), C.c_float(K["fx"]), C.c_float(K["fy"]),
                                    C.c_float(K["cx"]), C.c_float(K["cy"]), _p(w), _p(v))
    return w[:len(z)], v[:len(z)]


def frame_is_in_frustum_batch(tcw, ow, K, bounds, n_levels, log_sf, world_pos, normal, min_inv, max_inv, max_raw, cos_limit=0.5):
    tcw = np.ascontiguousarray(tcw, np.float32).reshape(-1, 12)
    ow = np.ascontiguousarray(ow, np.float32).reshape(-1, 3)
    wp = np.ascontiguousarray(world_pos, np.float32).reshape(-1, 3)
    no = np.ascontiguousarray(normal, np.float32).reshape(-1, 3)
    mi, ma, mr = (np.ascontiguousarray(a, np.float32) for a in (min_inv, max_inv, max_raw))
    n, m = len(tcw), len(wp)
    b = np.asarray(bounds, np.float32)
    iv = np.zeros((n, m), np.uint8)
    px, py, pxr, vc = (np.zeros((n, m), np.float32) for _ in range(4))
    lv = np.zeros((n, m), np.int32)
    lib().orc_frame_is_in_frustum_batch(C.c_int(n), _p(tcw), _p(ow), C.c_float(K["fx"]), C.c_float(K["fy"]), C.c_float(K["cx"]), C.c_float(K["cy"]),
                                        C.c_float(K["bf"]), _p(b), C.c_int(n_levels), C.c_float(log_sf), C.c_int(m), _p(wp), _p(no), _p(mi), _p(ma),
                                        _p(mr), C.c_float(cos_limit), _p(iv), _p(px), _p(py), _p(pxr), _p(lv), _p(vc))
    return iv, px, py, pxr, lv, vc


class _MapPointView(C.Structure):
    """layout of pl_mappoint_view (include/plslam_c.h)"""
    _fields_ = [("n", C.c_int), ("desc", C.c_void_p), ("track_in_view", C.c_void_p), ("proj_x", C.c_void_p), ("proj_y", C.c_void_p),
                ("proj_xr", C.c_void_p), ("scale_level", C.c_void_p), ("view_cos", C.c_void_p), ("has_observations", C.c_void_p)]


def _f32_at(addr, count):
    return np.ctypeslib.as_array((C.c_float * count).from_address(addr)) if count else np.zeros(0, np.float32)


def search_local_map(frame_view, ow, map_view, cos_limit, log_sf, th, nn_ratio):
    """Tracking::SearchLocalPoints (Tracking.cc:1746-1813) for one frame: Frame::isInFrustum of every map point of the snapshot, then
    ORBmatcher::SearchByProjection(F, localPoints, th) — the oracle twin of pl_orb_search_local_map_batch.
    frame_view / map_view: pl_frame_view / pl_localmap_view structures -> (match_of_feature, nmatches, map points in view)."""
    m = map_view.n
    K = dict(fx=frame_view.fx, fy=frame_view.fy, cx=frame_view.cx, cy=frame_view.cy, bf=frame_view.bf)
    tcw = np.array(list(frame_view.tcw), np.float32)
    bounds = (frame_view.min_x, frame_view.min_y, frame_view.max_x, frame_view.max_y)
    iv, px, py, pxr, lv, vc = frame_is_in_frustum_batch(tcw[None], np.asarray(ow, np.float32).reshape(1, 3), K, bounds, frame_view.n_levels, log_sf,
                                                        _f32_at(map_view.world_pos, 3 * m), _f32_at(map_view.normal, 3 * m),
                                                        _f32_at(map_view.min_dist_inv, m), _f32_at(map_view.max_dist_inv, m),
                                                        _f32_at(map_view.max_dist, m), cos_limit)
    mv = _MapPointView()
    mv.n = m
    mv.desc, mv.has_observations = map_view.desc, map_view.has_observations
    mv.track_in_view, mv.proj_x, mv.proj_y, mv.proj_xr = iv.ctypes.data, px.ctypes.data, py.ctypes.data, pxr.ctypes.data
    mv.scale_level, mv.view_cos = lv.ctypes.data, vc.ctypes.data
    match, n = search_local_points(frame_view, mv, th, nn_ratio)
    return match, n, int(iv.sum())


def frame_lines_in_frustum_batch(tcw, start3d, end3d):
    tcw = np.ascontiguousarray(tcw, np.float32).reshape(-1, 12)
    s3 = np.ascontiguousarray(start3d, np.float64).reshape(-1, 3)
    e3 = np.ascontiguousarray(end3d, np.float64).reshape(-1, 3)
    iv = np.zeros((len(tcw), len(s3)), np.uint8)
    lib().orc_frame_lines_in_frustum_batch(C.c_int(len(tcw)), _p(tcw), C.c_int(len(s3)), _p(s3), _p(e3), _p(iv))
    return iv


def frame_compute_stereo_matches(left, right, keys_left, desc_left, keys_right, desc_right, bf, b):
    """left / right: OrbOracle objects after extract() on the rectified pair."""
    kl = np.ascontiguousarray(keys_left, KP_DTYPE)
    kr = np.ascontiguousarray(keys_right, KP_DTYPE)
    dl, dr = _rows(desc_left), _rows(desc_right)
    ur = np.empty(max(len(kl), 1), np.float32)
    d = np.empty(max(len(kl), 1), np.float32)
    lib().orc_frame_compute_stereo_matches(left._h, right._h, _p(kl), _p(dl), C.c_int(len(kl)), _p(kr), _p(dr), C.c_int(len(kr)),
                                           C.c_float(bf), C.c_float(b), _p(ur), _p(d))
    return ur[:len(kl)], d[:len(kl)]


def frame_undistort_keylines(kls, K, dist_coef, img_size):
    kls = np.ascontiguousarray(kls, KL_DTYPE)
    dc = np.ascontiguousarray(dist_coef, np.float32).reshape(-1)[:5].copy()
    out = np.empty_like(kls)
    lib().orc_frame_undistort_keylines(_p(kls), C.c_int(len(kls)), C.c_float(K["fx"]), C.c_float(K["fy"]), C.c_float(K["cx"]), C.c_float(K["cy"]), _p(dc),
                                       C.c_int(int(img_size[0])), C.c_int(int(img_size[1])), _p(out))
    return out


def frame_assign_features_to_grid(keys_un, bounds):
    k = np.ascontiguousarray(keys_un, KP_DTYPE)
    b = np.asarray(bounds, np.float32)
    cst = np.zeros(64 * 48 + 1, np.int32)
    idx = np.zeros(max(len(k), 1), np.int32)
    lib().orc_frame_assign_features_to_grid(_p(k), C.c_int(len(k)), _p(b), _p(cst), _p(idx))
    return cst, idx[:cst[-1]]


class VocOracle:
    """CPU restatement of DBoW2's vocabulary transform (bow_oracle.cpp)."""

    def __init__(self):
        self._h = None
        lib().orc_voc_create.restype = C.c_void_p
        lib().orc_voc_load_text.restype = C.c_void_p
        lib().orc_voc_destroy.argtypes = [C.c_void_p]

    def __del__(self):
        if self._h:
            lib().orc_voc_destroy(self._h)
            self._h = None

    def create(self, k, L, scoring, weighting, parent, is_leaf, desc, weight):
        parent = np.ascontiguousarray(parent, np.int32)
        leaf = np.ascontiguousarray(is_leaf, np.uint8)
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        weight = np.ascontiguousarray(weight, np.float64)
        self._h = C.c_void_p(lib().orc_voc_create(C.c_int(k), C.c_int(L), C.c_int(scoring), C.c_int(weighting), C.c_int(len(parent)), _p(parent), _p(leaf),
                                                  _p(desc), _p(weight)))
        assert self._h
        return self

    def loadFromTextFile(self, filename):
        h = lib().orc_voc_load_text(str(filename).encode())
        self._h = C.c_void_p(h) if h else None
        return bool(h)

    def info(self):
        k, L, nn, nw = (C.c_int() for _ in range(4))
        lib().orc_voc_info(self._h, C.byref(k), C.byref(L), C.byref(nn), C.byref(nw))
        return dict(k=k.value, L=L.value, n_nodes=nn.value, n_words=nw.value)

    def transform(self, desc, levelsup=4):
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(d)
        cap = max(n, 1)
        wid, nid, fi = (np.zeros(cap, np.uint32) for _ in range(3))
        wv = np.zeros(cap, np.float64)
        noff = np.zeros(cap + 1, np.int32)
        nw, nn = C.c_int(), C.c_int()
        lib().orc_voc_transform(self._h, _p(d), C.c_int(n), C.c_int(levelsup), C.byref(nw), _p(wid), _p(wv), C.byref(nn), _p(nid), _p(noff), _p(fi))
        fv = {int(nid[k]): [int(x) for x in fi[noff[k]:noff[k + 1]]] for k in range(nn.value)}
        return (wid[:nw.value].copy(), wv[:nw.value].copy()), fv


class OracleBackend:
    """CPU-oracle backend for frontend.TrackingFrontEnd (same interface as frontend.GpuBackend)."""

    def __init__(self, nfeatures=1000, threads=1):
        self.orb = OrbOracle(nfeatures)
        self.threads = threads

    def scale_factors(self):
        return self.orb.tables()["scale_factors"]

    def extract_orb(self, frames):
        return [self.orb.extract(f) for f in frames]

    def extract_lines(self, frames):
        return [line_extract(f, 80) for f in frames]

    def search_last_frame(self, cv, lv, th):
        return search_last_frame(cv, lv, th)

    def search_local_points(self, fv, mv, th, nn):
        return search_local_points(fv, mv, th, nn)

    def project_lines(self, *a):
        return project_lines(*a)

    def match_lines(self, *a):
        return match_lines(*a)

    # "batched" forms: the CPU path simply loops (sequential matching, as in the reference)
    def search_last_frame_batch(self, cvs, lvs, th):
        return [search_last_frame(c, l, th) for c, l in zip(cvs, lvs)]

    def search_local_points_batch(self, fvs, mvs, th, nn):
        return [search_local_points(f, m, th, nn) for f, m in zip(fvs, mvs)]

    def search_local_map_batch(self, fvs, ow, maps, map_of_frame, cos_limit, log_sf, th, nn):
        return [search_local_map(f, ow[i], maps[map_of_frame[i]], cos_limit, log_sf, th, nn) for i, f in enumerate(fvs)]

    def line_search_batch(self, cvs, lvs):
        return [line_search_by_projection(c, l) for c, l in zip(cvs, lvs)]

    def unproject_batch(self, off, xy, z, rwc, ow, K):
        return frame_unproject_batch(off, xy, z, rwc, ow, K)

    def is_in_frustum_batch(self, *a):
        return frame_is_in_frustum_batch(*a)
