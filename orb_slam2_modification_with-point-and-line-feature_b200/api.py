"""Python mirrors of the reference's front-end classes, on top of the C ABI (include/plslam_c.h).

Names, argument meaning and error behaviour follow the reference:

* :class:`ORBextractor`  — reference include/ORBextractor.h:44-107 (ctor, ``operator()``, the getters, the public
  ``mvImagePyramid``).  An empty image returns silently with no keypoints (src/ORBextractor.cc:1046-1047); a
  non-uint8 / non-2-D image raises (the reference asserts, :1050).
* :class:`LineExtractor` — include/LineExtractor.h:23-30.
* :class:`DescriptorMatcher` — the Hamming searches of ORBmatcher / LineMatcher.

These classes are plumbing for tests and bench.py; the product is the native library.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N
from ._native import KL_DTYPE, KP_DTYPE, PlError, check, ptr


def _as_gray(image):
    a = np.asarray(image)
    if a.size == 0:
        return None
    if a.dtype != np.uint8 or a.ndim != 2:
        raise AssertionError("image.type() == CV_8UC1")  # ORBextractor.cc:1050
    if a.strides[1] != 1:
        a = np.ascontiguousarray(a)
    return a


class ORBextractor:
    """ORB_SLAM2::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)."""

    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7, device=0, max_cols=640,
                 max_rows=480, max_batch=1):
        self._h = C.c_void_p()
        self.nlevels = nlevels
        check(N.lib().pl_orb_create(C.byref(self._h), C.c_int(nfeatures), C.c_float(scaleFactor), C.c_int(nlevels),
                                    C.c_int(iniThFAST), C.c_int(minThFAST), C.c_int(device), C.c_int(max_cols),
                                    C.c_int(max_rows), C.c_int(max_batch)))
        self.max_batch = max_batch
        self._last = None  # (n_frames, rows, cols)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            N.lib().pl_orb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- getters (ORBextractor.h:63-84) ----
    def GetLevels(self):
        return N.lib().pl_orb_levels(self._h)

    def GetScaleFactor(self):
        return float(N.lib().pl_orb_scale_factor(self._h))

    def _vec(self, fn, dtype=np.float32):
        out = np.empty(self.nlevels, dtype)
        check(fn(self._h, ptr(out)))
        return out

    def GetScaleFactors(self):
        return self._vec(N.lib().pl_orb_scale_factors)

    def GetInverseScaleFactors(self):
        return self._vec(N.lib().pl_orb_inv_scale_factors)

    def GetScaleSigmaSquares(self):
        return self._vec(N.lib().pl_orb_level_sigma2)

    def GetInverseScaleSigmaSquares(self):
        return self._vec(N.lib().pl_orb_inv_level_sigma2)

    def features_per_level(self):
        return self._vec(N.lib().pl_orb_features_per_level, np.int32)

    def max_keypoints(self):
        return N.lib().pl_orb_max_keypoints(self._h)

    def last_launches(self):
        return N.lib().pl_orb_last_launches(self._h)

    # ---- operator() ----
    def __call__(self, image, mask=None):
        """Returns (keypoints[KP_DTYPE], descriptors uint8 (n,32)).  `mask` is ignored, as in the reference."""
        img = _as_gray(image)
        if img is None:
            return np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        cap = self.max_keypoints()
        kps = np.empty(cap, KP_DTYPE)
        desc = np.empty((cap, 32), np.uint8)
        n = C.c_int(0)
        check(N.lib().pl_orb_extract(self._h, ptr(img), C.c_int(img.shape[0]), C.c_int(img.shape[1]),
                                     C.c_size_t(img.strides[0]), ptr(kps), ptr(desc), C.c_int(cap), C.byref(n)))
        self._last = (1, img.shape[0], img.shape[1])
        return kps[:n.value].copy(), desc[:n.value].copy()

    def extract_batch(self, frames):
        """frames: uint8 (n, rows, cols) C-contiguous.  Returns (kps (n,cap), desc (n,cap,32), counts (n,))."""
        fr = np.asarray(frames)
        assert fr.dtype == np.uint8 and fr.ndim == 3 and fr.strides[2] == 1
        n, rows, cols = fr.shape
        cap = self.max_keypoints()
        kps = np.zeros((n, cap), KP_DTYPE)
        desc = np.zeros((n, cap, 32), np.uint8)
        cnt = np.zeros(n, np.int32)
        check(N.lib().pl_orb_extract_batch(self._h, ptr(fr), C.c_int(n), C.c_int(rows), C.c_int(cols),
                                           C.c_size_t(fr.strides[1]), C.c_size_t(fr.strides[0]), ptr(kps), ptr(desc),
                                           C.c_int(cap), ptr(cnt)))
        self._last = (min(n, self.max_batch) if n % self.max_batch == 0 else n % self.max_batch, rows, cols)
        return kps, desc, cnt

    def extract_batch_into(self, frames, kps, desc, cnt):
        """The host-pointer entry (pl_orb_extract_batch) with caller-owned outputs: kps (n, cap) KP_DTYPE, desc (n, cap, 32) uint8,
        cnt (n,) int32 — e.g. views of pinned memory, so that the copies inside the call are asynchronous DMA."""
        fr = np.asarray(frames)
        assert fr.dtype == np.uint8 and fr.ndim == 3 and fr.strides[2] == 1
        n, rows, cols = fr.shape
        cap = kps.shape[1]
        assert kps.shape[0] >= n and desc.shape[:2] == kps.shape[:2] and len(cnt) >= n
        check(N.lib().pl_orb_extract_batch(self._h, ptr(fr), C.c_int(n), C.c_int(rows), C.c_int(cols),
                                           C.c_size_t(fr.strides[1]), C.c_size_t(fr.strides[0]), ptr(kps), ptr(desc),
                                           C.c_int(cap), ptr(cnt)))
        self._last = (min(n, self.max_batch) if n % self.max_batch == 0 else n % self.max_batch, rows, cols)

    def stage_batch(self, frames):
        """pl_orb_stage_batch: enqueue the upload of one chunk of host frames and return (extract_staged_into does the rest)."""
        fr = np.asarray(frames)
        assert fr.dtype == np.uint8 and fr.ndim == 3 and fr.strides[2] == 1
        n, rows, cols = fr.shape
        check(N.lib().pl_orb_stage_batch(self._h, ptr(fr), C.c_int(n), C.c_int(rows), C.c_int(cols), C.c_size_t(fr.strides[1]), C.c_size_t(fr.strides[0])))
        self._keep_staged = fr   # the copy is asynchronous: the host frames stay alive until extract_staged_into returns
        self._last = (n, rows, cols)

    def stream_wait_staged(self, stream):
        """pl_orb_stream_wait_staged: `stream` (a raw cudaStream_t, e.g. LineExtractor.stream()) waits for the staging copy."""
        check(N.lib().pl_orb_stream_wait_staged(self._h, C.c_void_p(stream)))

    def extract_staged_into(self, kps, desc, cnt):
        """pl_orb_extract_staged with caller-owned outputs (as extract_batch_into)."""
        cap = kps.shape[1]
        assert desc.shape[:2] == kps.shape[:2] and len(cnt) >= self._last[0] and kps.shape[0] >= self._last[0]
        check(N.lib().pl_orb_extract_staged(self._h, ptr(kps), ptr(desc), C.c_int(cap), ptr(cnt)))
        self._keep_staged = None

    def staged_images(self):
        """(device address, n_frames, rows, cols, step, frame_stride) of the images the last host-pointer extract call left in HBM."""
        d, n, r, c_, st, fs = C.c_void_p(), C.c_int(), C.c_int(), C.c_int(), C.c_size_t(), C.c_size_t()
        check(N.lib().pl_orb_staged_images_dev(self._h, C.byref(d), C.byref(n), C.byref(r), C.byref(c_), C.byref(st), C.byref(fs)))
        return d.value, n.value, r.value, c_.value, st.value, fs.value

    def extract_batch_dev(self, d_gray, n, rows, cols, step, frame_stride, d_kps, d_desc, cap, d_nout):
        """All pointers are raw device addresses (ints); asynchronous on the handle's stream."""
        check(N.lib().pl_orb_extract_batch_dev(self._h, ptr(d_gray), C.c_int(n), C.c_int(rows), C.c_int(cols),
                                               C.c_size_t(step), C.c_size_t(frame_stride), ptr(d_kps), ptr(d_desc),
                                               C.c_int(cap), ptr(d_nout)))

    def sync(self):
        check(N.lib().pl_orb_sync(self._h))

    def stream(self):
        return N.lib().pl_orb_stream(self._h)

    # ---- mvImagePyramid (ORBextractor.h:85) and test hooks ----
    def pyramid_level(self, level, frame=0, bordered=False):
        rows, cols = C.c_int(), C.c_int()
        check(N.lib().pl_orb_pyramid_dims(self._h, C.c_int(level), C.byref(rows), C.byref(cols)))
        out = np.empty((rows.value + 38, cols.value + 38), np.uint8)
        check(N.lib().pl_orb_pyramid_read(self._h, C.c_int(frame), C.c_int(level), ptr(out), C.c_size_t(out.strides[0])))
        return out if bordered else out[19:-19, 19:-19]

    @property
    def mvImagePyramid(self):
        return [self.pyramid_level(l) for l in range(self.nlevels)]

    def blurred_level(self, level, frame=0):
        rows, cols = C.c_int(), C.c_int()
        check(N.lib().pl_orb_pyramid_dims(self._h, C.c_int(level), C.byref(rows), C.byref(cols)))
        out = np.empty((rows.value, cols.value), np.uint8)
        check(N.lib().pl_orb_blurred_read(self._h, C.c_int(frame), C.c_int(level), ptr(out), C.c_size_t(out.strides[0])))
        return out

    def candidates(self, level, frame=0):
        cap = 1 << 16
        xs, ys, rs = (np.empty(cap, np.float32) for _ in range(3))
        n = C.c_int(0)
        check(N.lib().pl_orb_candidates_read(self._h, C.c_int(frame), C.c_int(level), ptr(xs), ptr(ys), ptr(rs),
                                             C.c_int(cap), C.byref(n)))
        return xs[:n.value].copy(), ys[:n.value].copy(), rs[:n.value].copy()


class LineExtractor:
    """ORB_SLAM2::LineExtractor — ExtractLineSegment(img, key_lines, line_descriptor, keyline_coefficients)."""

    NUMS_LINE_FEATURE = 80  # LineExtractor.cpp:23

    def __init__(self, device=0, max_cols=640, max_rows=480, max_batch=1):
        self._h = C.c_void_p()
        check(N.lib().pl_line_create(C.byref(self._h), C.c_int(device), C.c_int(max_cols), C.c_int(max_rows), C.c_int(max_batch)))
        self.max_batch = max_batch

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            N.lib().pl_line_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def ExtractLineSegment(self, img, max_lines=NUMS_LINE_FEATURE):
        """Returns (key_lines[KL_DTYPE], line_descriptor uint8 (n,32), keyline_coefficients float64 (n,3))."""
        im = _as_gray(img)
        if im is None:
            return np.zeros(0, KL_DTYPE), np.zeros((0, 32), np.uint8), np.zeros((0, 3), np.float64)
        kls = np.zeros(max_lines, KL_DTYPE)
        desc = np.zeros((max_lines, 32), np.uint8)
        co = np.zeros((max_lines, 3), np.float64)
        n = C.c_int(0)
        check(N.lib().pl_line_extract(self._h, ptr(im), C.c_int(im.shape[0]), C.c_int(im.shape[1]), C.c_size_t(im.strides[0]),
                                      C.c_int(max_lines), ptr(kls), ptr(desc), ptr(co), C.byref(n)))
        return kls[:n.value].copy(), desc[:n.value].copy(), co[:n.value].copy()

    def extract_batch(self, frames, max_lines=NUMS_LINE_FEATURE):
        fr = np.asarray(frames)
        assert fr.dtype == np.uint8 and fr.ndim == 3 and fr.strides[2] == 1
        n, rows, cols = fr.shape
        kls = np.zeros((n, max_lines), KL_DTYPE)
        desc = np.zeros((n, max_lines, 32), np.uint8)
        co = np.zeros((n, max_lines, 3), np.float64)
        cnt = np.zeros(n, np.int32)
        check(N.lib().pl_line_extract_batch(self._h, ptr(fr), C.c_int(n), C.c_int(rows), C.c_int(cols), C.c_size_t(fr.strides[1]),
                                            C.c_size_t(fr.strides[0]), C.c_int(max_lines), ptr(kls), ptr(desc), ptr(co), ptr(cnt)))
        return kls, desc, co, cnt

    def extract_batch_into(self, frames, max_lines, kls, desc, co, cnt):
        """pl_line_extract_batch with caller-owned outputs (kls (n, max_lines) KL_DTYPE, desc (n, max_lines, 32), co (n, max_lines, 3)
        float64, cnt (n,) int32)."""
        fr = np.asarray(frames)
        assert fr.dtype == np.uint8 and fr.ndim == 3 and fr.strides[2] == 1
        n, rows, cols = fr.shape
        assert kls.shape == (n, max_lines) and desc.shape == (n, max_lines, 32) and co.shape == (n, max_lines, 3) and len(cnt) >= n
        check(N.lib().pl_line_extract_batch(self._h, ptr(fr), C.c_int(n), C.c_int(rows), C.c_int(cols), C.c_size_t(fr.strides[1]),
                                            C.c_size_t(fr.strides[0]), C.c_int(max_lines), ptr(kls), ptr(desc), ptr(co), ptr(cnt)))

    def set_graph(self, on=True):
        """Small chunks (tracking mode) as one CUDA graph launch, re-used while the call's pointers and sizes stay the same."""
        check(N.lib().pl_line_set_graph(self._h, C.c_int(int(on))))

    def stream_wait_grow_start(self, stream):
        """`stream` (raw cudaStream_t) waits until the region grower of the extraction last enqueued here has been launched."""
        check(N.lib().pl_line_stream_wait_grow_start(self._h, C.c_void_p(stream)))

    def extract_batch_from_dev_into(self, d_gray, n, rows, cols, step, frame_stride, max_lines, kls, desc, co, cnt):
        """pl_line_extract_batch_from_dev: images already in HBM (e.g. ORBextractor.staged_images()), caller-owned host outputs."""
        assert kls.shape == (n, max_lines) and desc.shape == (n, max_lines, 32) and co.shape == (n, max_lines, 3) and len(cnt) >= n
        check(N.lib().pl_line_extract_batch_from_dev(self._h, ptr(d_gray), C.c_int(n), C.c_int(rows), C.c_int(cols), C.c_size_t(step),
                                                     C.c_size_t(frame_stride), C.c_int(max_lines), ptr(kls), ptr(desc), ptr(co), ptr(cnt)))

    def extract_batch_dev(self, d_gray, n, rows, cols, step, frame_stride, max_lines, d_kls, d_desc, d_coef, d_nout):
        check(N.lib().pl_line_extract_batch_dev(self._h, ptr(d_gray), C.c_int(n), C.c_int(rows), C.c_int(cols), C.c_size_t(step),
                                                C.c_size_t(frame_stride), C.c_int(max_lines), ptr(d_kls), ptr(d_desc), ptr(d_coef),
                                                ptr(d_nout)))

    def sync(self):
        check(N.lib().pl_line_sync(self._h))

    def stream(self):
        return N.lib().pl_line_stream(self._h)

    def last_launches(self):
        return N.lib().pl_line_last_launches(self._h)

    def set_reserved_sms(self, n):
        """SMs the persistent region grower leaves to the kernels of other streams (matchers, Frame glue) in large batches."""
        check(N.lib().pl_line_set_reserved_sms(self._h, C.c_int(int(n))))

    # ---- test hooks ----
    def lsd_segments(self, frame=0, cap=30000):
        xy = np.empty((cap, 4), np.float32)
        w, p, nf = (np.empty(cap, np.float64) for _ in range(3))
        n = C.c_int(0)
        check(N.lib().pl_line_lsd_read(self._h, C.c_int(frame), ptr(xy), ptr(w), ptr(p), ptr(nf), C.c_int(cap), C.byref(n)))
        return xy[:n.value].copy(), w[:n.value].copy(), p[:n.value].copy(), nf[:n.value].copy()

    def scaled_image(self, frame=0):
        r, c = C.c_int(), C.c_int()
        check(N.lib().pl_line_scaled_dims(self._h, C.byref(r), C.byref(c)))
        out = np.empty((r.value, c.value), np.uint8)
        check(N.lib().pl_line_scaled_read(self._h, C.c_int(frame), ptr(out), C.c_size_t(out.strides[0])))
        return out

    def angle_map(self, frame=0):
        r, c = C.c_int(), C.c_int()
        check(N.lib().pl_line_scaled_dims(self._h, C.byref(r), C.byref(c)))
        out = np.empty((r.value, c.value), np.float32)
        check(N.lib().pl_line_angles_read(self._h, C.c_int(frame), ptr(out)))
        return out

    def float_descriptors(self, n, frame=0):
        out = np.empty((n, 72), np.float32)
        check(N.lib().pl_line_fdesc_read(self._h, C.c_int(frame), ptr(out), C.c_int(n)))
        return out


class _DenseRows:
    """Sequence of per-search results (row of matches, scalar, scalar, ...) backed by dense arrays: `dense` = (rows (n, widest) —
    a row is valid up to its own length —, the scalar columns...).  Stands in for the list of tuples the batched searches returned,
    without building n small arrays and n tuples when the caller reads the dense arrays."""

    def __init__(self, out, nl, *cols):
        self.dense = (out,) + cols
        self._nl = nl

    def __len__(self):
        return len(self._nl)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        return (self.dense[0][i, :self._nl[i]],) + tuple(int(c[i]) for c in self.dense[1:])

    def __iter__(self):
        return (self[i] for i in range(len(self)))


_LineMatches = _DenseRows   # (matches padded with -1, nmatches, used_relaxed, n_projected)


def _dense_out(views, typ, fill=None):
    """One (n, widest view) int32 array for the per-view outputs of a batched search + the array of row pointers the C ABI takes."""
    n = len(views)
    nl = np.frombuffer(views, np.dtype(typ), count=n)["n"].astype(np.int64) if n else np.zeros(0, np.int64)
    stride = max(int(nl.max()), 1) if n else 1
    out = np.empty((max(n, 1), stride), np.int32) if fill is None else np.full((max(n, 1), stride), fill, np.int32)
    ptrs = (out.ctypes.data + np.arange(max(n, 1), dtype=np.uint64) * np.uint64(stride * 4)).astype(np.uint64)
    return out, nl, ptrs


class DescriptorMatcher:
    """Hamming searches shared by ORBmatcher (src/ORBmatcher.cc) and LineMatcher (src/LineMatcher.cpp)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        check(N.lib().pl_match_create(C.byref(self._h), C.c_int(device)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            N.lib().pl_match_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(N.lib().pl_match_sync(self._h))

    def stream(self):
        return N.lib().pl_match_stream(self._h)

    def last_launches(self):
        return N.lib().pl_match_last_launches(self._h)

    @staticmethod
    def _rows(a):
        a = np.ascontiguousarray(a, np.uint8)
        assert a.ndim == 2 and a.shape[1] == 32
        return a

    def DescriptorDistance(self, a, b):
        """ORBmatcher::DescriptorDistance / LineMatcher::DescriptorDistance for row pairs."""
        a, b = self._rows(a), self._rows(b)
        assert a.shape == b.shape
        out = np.empty(a.shape[0], np.int32)
        check(N.lib().pl_hamming_pairs(self._h, ptr(a), ptr(b), C.c_int(a.shape[0]), ptr(out)))
        return out

    def knnMatch2(self, query, train):
        """cv::BFMatcher(NORM_HAMMING).knnMatch(query, train, 2) -> (idx (nq,2), dist (nq,2))."""
        q, t = self._rows(query), self._rows(train)
        idx = np.empty((q.shape[0], 2), np.int32)
        dist = np.empty((q.shape[0], 2), np.int32)
        check(N.lib().pl_hamming_knn2(self._h, ptr(q), C.c_int(q.shape[0]), ptr(t) if t.shape[0] else ptr(None),
                                      C.c_int(t.shape[0]), ptr(idx), ptr(dist)))
        return idx, dist

    def knn2_dev(self, d_q, nq, d_t, nt, d_idx, d_dist):
        check(N.lib().pl_hamming_knn2_dev(self._h, ptr(d_q), C.c_int(nq), ptr(d_t), C.c_int(nt), ptr(d_idx), ptr(d_dist)))

    def candidate_distances(self, query, train, cand_off, cand_idx):
        q, t = self._rows(query), self._rows(train)
        off = np.ascontiguousarray(cand_off, np.int32)
        ci = np.ascontiguousarray(cand_idx, np.int32)
        assert off.shape[0] == q.shape[0] + 1
        out = np.empty(ci.shape[0], np.int32)
        check(N.lib().pl_hamming_candidates(self._h, ptr(q), C.c_int(q.shape[0]), ptr(t), C.c_int(t.shape[0]), ptr(off),
                                            ptr(ci), ptr(out)))
        return out

    # ---- ORBmatcher projection searches (views built with _native.make_*_view) ----
    def SearchByProjectionLastFrame(self, cur_view, last_view, th, mono=False, check_orientation=True):
        """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) -> (match_of_feature, nmatches)."""
        match = np.empty(max(cur_view.n, 1), np.int32)
        n = C.c_int(0)
        check(N.lib().pl_orb_search_last_frame(self._h, C.byref(cur_view), C.byref(last_view), C.c_float(th), C.c_int(int(mono)),
                                               C.c_int(int(check_orientation)), ptr(match), C.byref(n)))
        return match[:cur_view.n], n.value

    def SearchByProjectionLocalPoints(self, frame_view, mp_view, th, nn_ratio):
        """ORBmatcher(nn_ratio).SearchByProjection(F, vpMapPoints, th) -> (match_of_feature, nmatches)."""
        match = np.empty(max(frame_view.n, 1), np.int32)
        n = C.c_int(0)
        check(N.lib().pl_orb_search_local_points(self._h, C.byref(frame_view), C.byref(mp_view), C.c_float(th), C.c_float(nn_ratio),
                                                 ptr(match), C.byref(n)))
        return match[:frame_view.n], n.value

    # ---- LineMatcher ----
    def project_lines(self, start3d, end3d, src_kl, valid, tcw, K, bounds, img_size):
        """Front half of LineMatcher::SearchByProjection -> (new_KeyLines, new_kl_index)."""
        s3 = np.ascontiguousarray(start3d, np.float64).reshape(-1, 3)
        e3 = np.ascontiguousarray(end3d, np.float64).reshape(-1, 3)
        kl = np.ascontiguousarray(src_kl, KL_DTYPE)
        va = np.ascontiguousarray(valid, np.uint8)
        n = len(kl)
        out = np.zeros(max(n, 1), KL_DTYPE)
        idx = np.zeros(max(n, 1), np.int32)
        m = C.c_int(0)
        t = (C.c_float * 12)(*[float(x) for x in np.asarray(tcw, np.float32).reshape(-1)[:12]])
        check(N.lib().pl_line_project(self._h, ptr(s3), ptr(e3), ptr(kl), ptr(va), C.c_int(n), t, C.c_float(K["fx"]), C.c_float(K["fy"]),
                                      C.c_float(K["cx"]), C.c_float(K["cy"]), C.c_float(bounds[0]), C.c_float(bounds[1]),
                                      C.c_float(bounds[2]), C.c_float(bounds[3]), C.c_int(img_size[0]), C.c_int(img_size[1]), ptr(out),
                                      ptr(idx), C.byref(m)))
        return out[:m.value].copy(), idx[:m.value].copy()

    def match_lines(self, proj_kl, proj_desc, cur_kl, cur_desc, cur_claimed=None):
        """Back half of LineMatcher::SearchByProjection -> (match_of_line, nmatches, used_relaxed)."""
        pk = np.ascontiguousarray(proj_kl, KL_DTYPE)
        pd = np.ascontiguousarray(proj_desc, np.uint8).reshape(-1, 32)
        ck = np.ascontiguousarray(cur_kl, KL_DTYPE)
        cd = np.ascontiguousarray(cur_desc, np.uint8).reshape(-1, 32)
        cc = None if cur_claimed is None else np.ascontiguousarray(cur_claimed, np.uint8)
        match = np.full(max(len(ck), 1), -1, np.int32)
        n, rel = C.c_int(0), C.c_int(0)
        check(N.lib().pl_line_match_pairs(self._h, ptr(pk), ptr(pd), C.c_int(len(pk)), ptr(ck), ptr(cd), ptr(cc), C.c_int(len(ck)),
                                          ptr(match), C.byref(n), C.byref(rel)))
        return match[:len(ck)], n.value, rel.value

    # ---- batched forms: n independent reference calls in one pass ----
    @staticmethod
    def _view_array(views, typ):
        if isinstance(views, C.Array) and views._type_ is typ:  # the caller built the array of views itself
            return views
        arr = (typ * len(views))()
        for i, v in enumerate(views):
            arr[i] = v
        return arr

    def _batch_points(self, fn, frame_views, pt_views, pt_type, *scalars):
        n = len(frame_views)
        fa = self._view_array(frame_views, N.FrameView)
        pa = self._view_array(pt_views, pt_type)
        out, nl, ptrs = _dense_out(fa, N.FrameView)
        cnt = np.zeros(max(n, 1), np.int32)
        check(fn(self._h, C.c_int(n), fa, pa, *scalars, ptr(ptrs), ptr(cnt)))
        return _DenseRows(out[:n], nl, cnt[:n]) if n else []

    def SearchByProjectionLastFrameBatch(self, cur_views, last_views, th, mono=False, check_orientation=True):
        return self._batch_points(N.lib().pl_orb_search_last_frame_batch, cur_views, last_views, N.LastFrameView, C.c_float(th),
                                  C.c_int(int(mono)), C.c_int(int(check_orientation)))

    def SearchByProjectionLocalPointsBatch(self, frame_views, mp_views, th, nn_ratio):
        return self._batch_points(N.lib().pl_orb_search_local_points_batch, frame_views, mp_views, N.MapPointView, C.c_float(th),
                                  C.c_float(nn_ratio))

    def SearchLocalMapBatch(self, frame_views, ow, map_views, map_of_frame, viewing_cos_limit, log_scale_factor, th, nn_ratio):
        """Tracking::SearchLocalPoints for n frames: Frame::IsInFrustum of the local map + ORBmatcher::SearchByProjection(F, localPoints)
        in one call, the projections staying on the device -> [(match_of_feature, nmatches, map points in view)]."""
        n = len(frame_views)
        fa = self._view_array(frame_views, N.FrameView)
        ma = self._view_array(map_views, N.LocalMapView)
        ow = np.ascontiguousarray(ow, np.float32).reshape(-1, 3)
        mof = np.ascontiguousarray(map_of_frame, np.int32).reshape(-1)
        assert ow.shape[0] == n and mof.shape[0] == n
        out, nl, ptrs = _dense_out(fa, N.FrameView)
        cnt, inv = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32)
        check(N.lib().pl_orb_search_local_map_batch(self._h, C.c_int(n), fa, ptr(ow), C.c_int(len(map_views)), ma, ptr(mof), C.c_float(viewing_cos_limit),
                                                    C.c_float(log_scale_factor), C.c_float(th), C.c_float(nn_ratio), ptr(ptrs), ptr(cnt), ptr(inv)))
        return _DenseRows(out[:n], nl, cnt[:n], inv[:n]) if n else []

    def SearchLinesByProjectionBatch(self, cur_views, line_views):
        """LineMatcher::SearchByProjection for n (frame, map lines) pairs ->
        [(match_of_line -> original map-line index or -1, nmatches, used_relaxed, n_projected)]."""
        n = len(cur_views)
        ca = self._view_array(cur_views, N.LineFrameView)
        la = self._view_array(line_views, N.MapLineView)
        # one (n, max lines) result array, rows padded with -1; `dense` on the returned sequence hands it over whole
        out, nl, ptrs = _dense_out(ca, N.LineFrameView, fill=-1)
        cnt, rel, npj = (np.zeros(max(n, 1), np.int32) for _ in range(3))
        nulls = (C.c_void_p * max(n, 1))()
        check(N.lib().pl_line_search_by_projection_batch(self._h, C.c_int(n), ca, la, ptr(ptrs), ptr(cnt), ptr(rel), nulls, nulls, ptr(npj)))
        return _LineMatches(out[:n], nl, cnt[:n], rel[:n], npj[:n]) if n else []

    # ---- C4 / C5: pose-based projection searches ----
    def _pose_points(self, fn, frame_views, pt_views, ow, log_sf, *scalars):
        n = len(frame_views)
        fa = self._view_array(frame_views, N.FrameView)
        pa = self._view_array(pt_views, N.PosePointView)
        ow = np.ascontiguousarray(ow, np.float32).reshape(-1, 3)
        ls = np.ascontiguousarray(log_sf, np.float32).reshape(-1)
        assert ow.shape[0] == n and ls.shape[0] == n
        outs = [np.empty(max(v.n, 1), np.int32) for v in frame_views]
        ptrs = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
        cnt = np.zeros(max(n, 1), np.int32)
        check(fn(self._h, C.c_int(n), fa, pa, ptr(ow), ptr(ls), *scalars, ptrs, ptr(cnt)))
        return [(outs[i][:frame_views[i].n], int(cnt[i])) for i in range(n)]

    def SearchByProjectionKeyFrameBatch(self, cur_views, pt_views, ow, log_scale_factor, th, orb_dist, check_orientation=True):
        """ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) for n calls ->
        [(match_of_feature -> key-frame feature index or -1, nmatches)]."""
        return self._pose_points(N.lib().pl_orb_search_keyframe_points_batch, cur_views, pt_views, ow, log_scale_factor, C.c_float(th),
                                 C.c_int(int(orb_dist)), C.c_int(int(check_orientation)))

    def SearchByProjectionSim3Batch(self, kf_views, pt_views, ow, log_scale_factor, th):
        """ORBmatcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th) for n calls ->
        [(match_of_feature -> index into vpPoints or -1, nmatches)]."""
        return self._pose_points(N.lib().pl_orb_search_sim3_points_batch, kf_views, pt_views, ow, log_scale_factor, C.c_int(int(th)))

    # ---- C6 / C7: SearchByBoW ----
    def SearchByBoWBatch(self, a_views, b_views, mode, nn_ratio, check_orientation=True):
        """mode 0: SearchByBoW(pKF, F, vpMapPointMatches) -> per call (match over F features -> KF feature, nmatches);
        mode 1: SearchByBoW(pKF1, pKF2, vpMatches12) -> per call (match over KF1 features -> KF2 feature, nmatches)."""
        n = len(a_views)
        aa = self._view_array(a_views, N.BowView)
        ba = self._view_array(b_views, N.BowView)
        sizes = [(b_views[i].n if mode == 0 else a_views[i].n) for i in range(n)]
        outs = [np.empty(max(sz, 1), np.int32) for sz in sizes]
        ptrs = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
        cnt = np.zeros(max(n, 1), np.int32)
        check(N.lib().pl_orb_search_bow_batch(self._h, C.c_int(n), aa, ba, C.c_int(mode), C.c_float(nn_ratio), C.c_int(int(check_orientation)),
                                              ptrs, ptr(cnt)))
        return [(outs[i][:sizes[i]], int(cnt[i])) for i in range(n)]

    # ---- D6: brute-force line matchers ----
    def LineMatchKnnRatio(self, ref_desc, cur_desc):
        """LineMatcher::SearchByProjection(CurrentFrame, RefFrame, vpMapLineMatches) -> (match_of_line, nmatches)."""
        r, c = self._rows(ref_desc), self._rows(cur_desc)
        match = np.full(max(c.shape[0], 1), -1, np.int32)
        n = C.c_int(0)
        check(N.lib().pl_line_match_knn_ratio(self._h, ptr(r), C.c_int(r.shape[0]), ptr(c), C.c_int(c.shape[0]), ptr(match), C.byref(n)))
        return match[:c.shape[0]], n.value

    def LineSearchForTriangulation(self, desc1, desc2):
        """LineMatcher::SearchForTriangulation -> (pairs [k,2], nn_mad, nn12_mad)."""
        a, b = self._rows(desc1), self._rows(desc2)
        pairs = np.zeros((max(a.shape[0], 1), 2), np.int32)
        n = C.c_int(0)
        m1, m2 = C.c_double(0), C.c_double(0)
        check(N.lib().pl_line_search_for_triangulation(self._h, ptr(a), C.c_int(a.shape[0]), ptr(b), C.c_int(b.shape[0]), ptr(pairs),
                                                       C.byref(n), C.byref(m1), C.byref(m2)))
        return pairs[:n.value].copy(), m1.value, m2.value

    def LineFuseCandidates(self, ml_desc, valid, kf_desc):
        """LineMatcher::Fuse, descriptor half -> (tdx per map line or -1, nFused)."""
        a, b = self._rows(ml_desc), self._rows(kf_desc)
        va = None if valid is None else np.ascontiguousarray(valid, np.uint8)
        tdx = np.full(max(a.shape[0], 1), -1, np.int32)
        n = C.c_int(0)
        check(N.lib().pl_line_fuse_candidates(self._h, ptr(a), ptr(va), C.c_int(a.shape[0]), ptr(b), C.c_int(b.shape[0]), ptr(tdx), C.byref(n)))
        return tdx[:a.shape[0]], n.value

    # ---- E rows: the remaining ORBmatcher entry points (Fuse, SearchBySim3, SearchForInitialization,
    #      SearchForTriangulation) and ComputeDistinctiveDescriptors ----
    def FuseCandidatesBatch(self, kf_views, pt_views, ow, log_scale_factor, inv_level_sigma2, th, variant=0):
        """ORBmatcher::Fuse(pKF, vpMapPoints, th) (variant 0) / Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) (variant 1) for n
        calls -> [(best_idx per map point (-1 = no fusion), best_dist, n_fused)]."""
        n = len(kf_views)
        fa = self._view_array(kf_views, N.FrameView)
        pa = self._view_array(pt_views, N.PosePointView)
        ow = np.ascontiguousarray(ow, np.float32).reshape(-1, 3)
        ls = np.ascontiguousarray(log_scale_factor, np.float32).reshape(-1)
        sig = None if inv_level_sigma2 is None else [np.ascontiguousarray(s_, np.float32) for s_ in inv_level_sigma2]
        sp = (C.c_void_p * n)(*[s_.ctypes.data for s_ in sig]) if sig is not None else None
        bi = [np.empty(max(v.n, 1), np.int32) for v in pt_views]
        bd = [np.empty(max(v.n, 1), np.int32) for v in pt_views]
        pi = (C.c_void_p * n)(*[o.ctypes.data for o in bi])
        pd = (C.c_void_p * n)(*[o.ctypes.data for o in bd])
        cnt = np.zeros(max(n, 1), np.int32)
        check(N.lib().pl_orb_fuse_candidates_batch(self._h, C.c_int(n), fa, pa, ptr(ow), ptr(ls), sp, C.c_float(th), C.c_int(variant), pi, pd,
                                                   ptr(cnt)))
        return [(bi[i][:pt_views[i].n], bd[i][:pt_views[i].n], int(cnt[i])) for i in range(n)]

    def SearchBySim3(self, kf1, kf2, pts1, pts2, t21, t12, log_sf1, log_sf2, th):
        """ORBmatcher::SearchBySim3 -> (match12 over KF1 features -> KF2 feature or -1, nFound)."""
        t21 = np.ascontiguousarray(t21, np.float32).reshape(-1)[:12].copy()
        t12 = np.ascontiguousarray(t12, np.float32).reshape(-1)[:12].copy()
        m = np.empty(max(kf1.n, 1), np.int32)
        nf = C.c_int(0)
        check(N.lib().pl_orb_search_by_sim3(self._h, C.byref(kf1), C.byref(kf2), C.byref(pts1), C.byref(pts2), ptr(t21), ptr(t12),
                                            C.c_float(log_sf1), C.c_float(log_sf2), C.c_float(th), ptr(m), C.byref(nf)))
        return m[:kf1.n], nf.value

    def SearchForInitialization(self, f1, f2, prev_matched, window_size, nn_ratio=0.9, check_orientation=True):
        """ORBmatcher::SearchForInitialization -> (vnMatches12, nmatches, updated vbPrevMatched)."""
        pm = np.array(prev_matched, np.float32).reshape(-1, 2).copy()
        m = np.empty(max(f1.n, 1), np.int32)
        nm = C.c_int(0)
        check(N.lib().pl_orb_search_for_initialization(self._h, C.byref(f1), C.byref(f2), ptr(pm), C.c_int(int(window_size)), C.c_float(nn_ratio),
                                                       C.c_int(int(check_orientation)), ptr(m), C.byref(nm)))
        return m[:f1.n], nm.value, pm

    def SearchForTriangulation(self, a, b, f12, cw1, kf2_tcw, K2, scale_factors2, level_sigma2_2, only_stereo=False, check_orientation=True):
        """ORBmatcher::SearchForTriangulation -> (vMatchedPairs as an (n, 2) array, nmatches)."""
        f12 = np.ascontiguousarray(f12, np.float32).reshape(-1)[:9].copy()
        cw = np.ascontiguousarray(cw1, np.float32).reshape(-1)[:3].copy()
        t2 = np.ascontiguousarray(kf2_tcw, np.float32).reshape(-1)[:12].copy()
        sf = np.ascontiguousarray(scale_factors2, np.float32)
        sg = np.ascontiguousarray(level_sigma2_2, np.float32)
        pairs = np.empty((max(a.bow.n, 1), 2), np.int32)
        nm = C.c_int(0)
        check(N.lib().pl_orb_search_for_triangulation(self._h, C.byref(a), C.byref(b), ptr(f12), ptr(cw), ptr(t2), C.c_float(K2["fx"]),
                                                      C.c_float(K2["fy"]), C.c_float(K2["cx"]), C.c_float(K2["cy"]), ptr(sf), ptr(sg),
                                                      C.c_int(len(sf)), C.c_int(int(only_stereo)), C.c_int(int(check_orientation)), ptr(pairs),
                                                      C.byref(nm)))
        return pairs[:max(nm.value, 0)], nm.value

    def DistinctiveDescriptors(self, desc, group_off):
        """MapPoint / MapLine::ComputeDistinctiveDescriptors for many map points at once -> best row per group (-1 = empty)."""
        d = self._rows(desc)
        off = np.ascontiguousarray(group_off, np.int32)
        best = np.empty(max(len(off) - 1, 1), np.int32)
        check(N.lib().pl_distinctive_descriptors(self._h, ptr(d), ptr(off), C.c_int(len(off) - 1), ptr(best)))
        return best[:len(off) - 1]

    # ---- F rows: per-feature maps of Frame between the extractors and the matchers ----
    def UndistortPoints(self, xy, K, dist_coef):
        """Frame::UndistortKeyPoints: cv::undistortPoints(mat, mat, mK, mDistCoef, cv::Mat(), mK) for (n, 2) float points."""
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        dc = np.ascontiguousarray(dist_coef, np.float32).reshape(-1)[:5].copy()
        out = np.empty_like(xy)
        check(N.lib().pl_frame_undistort_points(self._h, ptr(xy), C.c_int(len(xy)), C.c_float(K["fx"]), C.c_float(K["fy"]), C.c_float(K["cx"]),
                                                C.c_float(K["cy"]), ptr(dc), ptr(out)))
        return out

    def UndistortKeyLines(self, kls, K, dist_coef, img_size):
        """Frame::UndistortKeyLines: end points through cv::undistortPoints, derived KeyLine fields recomputed."""
        kls = np.ascontiguousarray(kls, N.KL_DTYPE)
        dc = np.ascontiguousarray(dist_coef, np.float32).reshape(-1)[:5].copy()
        out = np.empty_like(kls)
        check(N.lib().pl_frame_undistort_keylines(self._h, ptr(kls), C.c_int(len(kls)), C.c_float(K["fx"]), C.c_float(K["fy"]), C.c_float(K["cx"]),
                                                  C.c_float(K["cy"]), ptr(dc), C.c_int(int(img_size[0])), C.c_int(int(img_size[1])), ptr(out)))
        return out

    def AssignFeaturesToGrid(self, keys_un, bounds):
        """Frame::AssignFeaturesToGrid -> (cell_start (64 * 48 + 1,), sorted_idx): cell x * 48 + y owns sorted_idx[cell_start[c]:cell_start[c + 1]]."""
        k = np.ascontiguousarray(keys_un, N.KP_DTYPE)
        b = np.asarray(bounds, np.float32)
        cst = np.zeros(64 * 48 + 1, np.int32)
        idx = np.zeros(max(len(k), 1), np.int32)
        check(N.lib().pl_frame_assign_features_to_grid(self._h, ptr(k), C.c_int(len(k)), ptr(b), ptr(cst), ptr(idx)))
        return cst, idx[:cst[-1]]

    def StereoFromRGBDBatch(self, depth, off, xy, x_un, bf, depth_dev_ptr=None):
        """Frame::ComputeStereoFromRGBD for the frames of a sequence -> (mvDepth, mvuRight) over the concatenated features.
        depth: (n_frames, rows, cols) float32 host array, or its shape when depth_dev_ptr gives a device copy."""
        off = np.ascontiguousarray(off, np.int32)
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        xu = np.ascontiguousarray(x_un, np.float32)
        if depth_dev_ptr is None:
            depth = np.ascontiguousarray(depth, np.float32)
            nf, rows, cols = depth.shape
            dp, isdev = ptr(depth), 0
        else:
            nf, rows, cols = depth
            dp, isdev = C.c_void_p(depth_dev_ptr), 1
        d = np.empty(max(len(xu), 1), np.float32)
        ur = np.empty(max(len(xu), 1), np.float32)
        check(N.lib().pl_frame_stereo_from_rgbd_batch(self._h, C.c_int(nf), dp, C.c_int(isdev), C.c_int(rows), C.c_int(cols), C.c_size_t(cols * 4),
                                                      C.c_size_t(rows * cols * 4), ptr(off), ptr(xy), ptr(xu), C.c_float(bf), ptr(d), ptr(ur)))
        return d[:len(xu)], ur[:len(xu)]

    def UnprojectBatch(self, off, xy_un, z, rwc, ow, K):
        """Frame::UnprojectStereo for the frames of a sequence -> (world (total, 3), valid (total,))."""
        off = np.ascontiguousarray(off, np.int32)
        xy = np.ascontiguousarray(xy_un, np.float32).reshape(-1, 2)
        z = np.ascontiguousarray(z, np.float32)
        rwc = np.ascontiguousarray(rwc, np.float32).reshape(-1, 9)
        ow = np.ascontiguousarray(ow, np.float32).reshape(-1, 3)
        w = np.empty((max(len(z), 1), 3), np.float32)
        v = np.empty(max(len(z), 1), np.uint8)
        check(N.lib().pl_frame_unproject_batch(self._h, C.c_int(len(off) - 1), ptr(off), ptr(xy), ptr(z), ptr(rwc), ptr(ow), C.c_float(K["fx"]),
                                               C.c_float(K["fy"]), C.c_float(K["cx"]), C.c_float(K["cy"]), ptr(w), ptr(v)))
        return w[:len(z)], v[:len(z)]

    def IsInFrustumBatch(self, tcw, ow, K, bounds, n_levels, log_sf, world_pos, normal, min_inv, max_inv, max_raw, cos_limit=0.5):
        """Frame::IsInFrustum of n frames against one snapshot of m map points -> (in_view, proj_x, proj_y, proj_xr, level, view_cos),
        each (n, m)."""
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
        check(N.lib().pl_frame_is_in_frustum_batch(self._h, C.c_int(n), ptr(tcw), ptr(ow), C.c_float(K["fx"]), C.c_float(K["fy"]), C.c_float(K["cx"]),
                                                   C.c_float(K["cy"]), C.c_float(K["bf"]), ptr(b), C.c_int(n_levels), C.c_float(log_sf), C.c_int(m),
                                                   ptr(wp), ptr(no), ptr(mi), ptr(ma), ptr(mr), C.c_float(cos_limit), ptr(iv), ptr(px), ptr(py),
                                                   ptr(pxr), ptr(lv), ptr(vc)))
        return iv, px, py, pxr, lv, vc

    def ComputeStereoMatches(self, left, right, keys_left, desc_left, keys_right, desc_right, bf, b, frame=0):
        """Frame::ComputeStereoMatches: left / right = the two ORBextractor objects after their extract call on the rectified pair
        -> (mvuRight, mvDepth)."""
        kl = np.ascontiguousarray(keys_left, N.KP_DTYPE)
        kr = np.ascontiguousarray(keys_right, N.KP_DTYPE)
        dl, dr = self._rows(desc_left), self._rows(desc_right)
        ur = np.empty(max(len(kl), 1), np.float32)
        d = np.empty(max(len(kl), 1), np.float32)
        check(N.lib().pl_frame_compute_stereo_matches(self._h, left._h, right._h, C.c_int(frame), ptr(kl), ptr(dl), C.c_int(len(kl)), ptr(kr), ptr(dr),
                                                      C.c_int(len(kr)), C.c_float(bf), C.c_float(b), ptr(ur), ptr(d)))
        return ur[:len(kl)], d[:len(kl)]

    def LinesInFrustumBatch(self, tcw, start3d, end3d):
        tcw = np.ascontiguousarray(tcw, np.float32).reshape(-1, 12)
        s3 = np.ascontiguousarray(start3d, np.float64).reshape(-1, 3)
        e3 = np.ascontiguousarray(end3d, np.float64).reshape(-1, 3)
        iv = np.zeros((len(tcw), len(s3)), np.uint8)
        check(N.lib().pl_frame_lines_in_frustum_batch(self._h, C.c_int(len(tcw)), ptr(tcw), C.c_int(len(s3)), ptr(s3), ptr(e3), ptr(iv)))
        return iv


class ORBVocabulary:
    """Mirror of ORB_SLAM2::ORBVocabulary (DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>) for the one call on the hot path:
    transform(features, BowVector, FeatureVector, levelsup) — Frame::ComputeBoW (Frame.cc:721-735)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        self._device = device

    def close(self):
        if getattr(self, "_h", None):
            N.lib().pl_voc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def create(self, k, L, scoring, weighting, parent, is_leaf, desc, weight):
        """The tree as loadFromTextFile builds it: node i+1 has parent[i] (0 = root), in file order."""
        self.close()
        parent = np.ascontiguousarray(parent, np.int32)
        leaf = np.ascontiguousarray(is_leaf, np.uint8)
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        weight = np.ascontiguousarray(weight, np.float64)
        check(N.lib().pl_voc_create(C.byref(self._h), C.c_int(self._device), C.c_int(k), C.c_int(L), C.c_int(scoring), C.c_int(weighting),
                                    C.c_int(len(parent)), ptr(parent), ptr(leaf), ptr(desc), ptr(weight)))
        return self

    def loadFromTextFile(self, filename):
        self.close()
        rc = N.lib().pl_voc_load_text(C.byref(self._h), C.c_int(self._device), str(filename).encode())
        if rc == N.PL_ERR_ARG:
            return False          # the reference returns false on a file it cannot parse
        check(rc)
        return True

    def info(self):
        k, L, nn, nw = (C.c_int() for _ in range(4))
        check(N.lib().pl_voc_info(self._h, C.byref(k), C.byref(L), C.byref(nn), C.byref(nw)))
        return dict(k=k.value, L=L.value, n_nodes=nn.value, n_words=nw.value)

    def transform_batch(self, descs, levelsup=4):
        """descs: list of (n_i, 32) uint8 arrays -> list of (BowVector as (word_id, word_value), FeatureVector as dict node -> [features])."""
        counts = [len(d) for d in descs]
        off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        tot, nf = int(off[-1]), len(descs)
        allv = np.concatenate([np.ascontiguousarray(d, np.uint8).reshape(-1, 32) for d in descs]) if tot else np.zeros((0, 32), np.uint8)
        cap = max(tot, 1)
        nw, nn = np.zeros(max(nf, 1), np.int32), np.zeros(max(nf, 1), np.int32)
        wid, nid, fi = (np.zeros(cap, np.uint32) for _ in range(3))
        wv = np.zeros(cap, np.float64)
        noff = np.zeros(cap + nf, np.int32)
        check(N.lib().pl_voc_transform_batch(self._h, C.c_int(nf), ptr(off), ptr(allv), C.c_int(levelsup), ptr(nw), ptr(wid), ptr(wv), ptr(nn), ptr(nid),
                                             ptr(noff), ptr(fi)))
        out = []
        for f in range(nf):
            o = int(off[f])
            no = noff[o + f:o + f + nn[f] + 1]
            fv = {int(nid[o + k]): [int(x) for x in fi[o + no[k]:o + no[k + 1]]] for k in range(nn[f])}
            out.append(((wid[o:o + nw[f]].copy(), wv[o:o + nw[f]].copy()), fv))
        return out

    def transform(self, desc, levelsup=4):
        return self.transform_batch([desc], levelsup)[0]
