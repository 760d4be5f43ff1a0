"""ctypes binding of oracle/_ref/libplref.so: the REFERENCE'S OWN hot-path code (src/ORBextractor.cc as a whole, DescriptorDistance
and ComputeThreeMaxima cut out of src/ORBmatcher.cc / src/LineMatcher.cpp, the two tracking searches
ORBmatcher::SearchByProjection (local map points, last frame) with the Frame grid functions they call, both ORBmatcher::SearchByBoW
overloads, the relocalisation and Sim3 searches, LineMatcher::SearchByProjection (last frame, local map) with LineMatching and
LiangBarsky, the vendored DBoW2 vocabulary), compiled from the sources
where they lie under /root/reference against the OpenCV stand-in of oracle/ref_shim/cv_standin.hpp (see oracle/ref_shim/Makefile).

TEST INFRASTRUCTURE: tests/test_oracle_ref.py checks the oracle's restatements against it.  The library can only be BUILT where
/root/reference exists; it is git-ignored but travels to the GPU box with the snapshot."""
import ctypes as C
import os
import subprocess

import numpy as np

import pyoracle

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libplref.so")
REFERENCE = os.environ.get("PLSLAM_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.exists(LIB_PATH) or os.path.isdir(os.path.join(REFERENCE, "src"))


def build(force: bool = False) -> str:
    """Builds libplref.so when the reference sources are present (no-op otherwise: the prebuilt file is used)."""
    if os.path.isdir(os.path.join(REFERENCE, "src")):
        pyoracle.build()
        args = ["make", "-C", os.path.join(_HERE, "ref_shim"), "-s", f"REF={REFERENCE}"]
        if force:
            args.append("-B")
        subprocess.check_call(args)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        pyoracle.lib()  # libplref.so resolves the image-processing primitives in libploracle.so
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.ref_voc_load_text.restype = C.c_void_p
        _lib.ref_voc_load_text.argtypes = [C.c_char_p]
        _lib.ref_voc_destroy.argtypes = [C.c_void_p]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def orb_extract(img, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, cap=20000, monotone=True):
    """ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)(img) -> (keypoints[KP_DTYPE], descriptors (n,32)).
    monotone: the reference's allocations come from a bump arena (a later node has the higher address), which makes the order
    DistributeOctTree gives nodes with equally many key points — the address, ORBextractor.cc:684 — the creation order."""
    img = np.ascontiguousarray(img, np.uint8)
    kps = np.zeros(cap, pyoracle.KP_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    n = C.c_int()
    rc = lib().ref_orb_extract(_p(img), C.c_int(img.shape[0]), C.c_int(img.shape[1]), C.c_size_t(img.strides[0]), C.c_int(nfeatures),
                               C.c_float(scale_factor), C.c_int(nlevels), C.c_int(ini_th), C.c_int(min_th), C.c_int(int(monotone)), _p(kps), _p(desc), C.c_int(cap),
                               C.byref(n))
    assert rc == 0
    return kps[:n.value].copy(), desc[:n.value].copy()


def distribute_octtree(xs, ys, resp, minX, maxX, minY, maxY, N, monotone=True):
    xs, ys, resp = (np.ascontiguousarray(a, np.float32) for a in (xs, ys, resp))
    cap = len(xs) + 8
    ox, oy, orr = (np.empty(cap, np.float32) for _ in range(3))
    n = lib().ref_distribute_octtree(_p(xs), _p(ys), _p(resp), C.c_int(len(xs)), C.c_int(minX), C.c_int(maxX), C.c_int(minY), C.c_int(maxY),
                                     C.c_int(N), C.c_int(int(monotone)), _p(ox), _p(oy), _p(orr), C.c_int(cap))
    return ox[:n].copy(), oy[:n].copy(), orr[:n].copy()


def orb_descriptor_distance(a, b):
    return lib().ref_orb_descriptor_distance(_p(np.ascontiguousarray(a, np.uint8)), _p(np.ascontiguousarray(b, np.uint8)))


def line_descriptor_distance(a, b):
    return lib().ref_line_descriptor_distance(_p(np.ascontiguousarray(a, np.uint8)), _p(np.ascontiguousarray(b, np.uint8)))


def compute_three_maxima(counts):
    counts = np.ascontiguousarray(counts, np.int32)
    ind = np.zeros(3, np.int32)
    lib().ref_compute_three_maxima(_p(counts), C.c_int(len(counts)), _p(ind))
    return tuple(int(v) for v in ind)


class Vocabulary:
    """ORBVocabulary = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>, loaded with loadFromTextFile."""

    def __init__(self, filename):
        self._h = lib().ref_voc_load_text(str(filename).encode())
        if not self._h:
            raise ValueError("loadFromTextFile failed: " + str(filename))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ref_voc_destroy(C.c_void_p(self._h))
            self._h = None

    def transform(self, desc, levelsup=4):
        """Frame::ComputeBoW: ((word ids, word values), {node id: feature indices})."""
        desc = np.ascontiguousarray(desc, np.uint8)
        n = len(desc)
        wid, wv = np.empty(n + 1, np.uint32), np.empty(n + 1, np.float64)
        nid, noff, fidx = np.empty(n + 1, np.uint32), np.empty(n + 2, np.int32), np.empty(n + 1, np.uint32)
        nw, nn = C.c_int(), C.c_int()
        lib().ref_voc_transform(C.c_void_p(self._h), _p(desc), C.c_int(n), C.c_int(levelsup), C.byref(nw), _p(wid), _p(wv), C.byref(nn), _p(nid),
                                _p(noff), _p(fidx))
        fv = {int(nid[i]): fidx[noff[i]:noff[i + 1]].copy() for i in range(nn.value)}
        return (wid[:nw.value].copy(), wv[:nw.value].copy()), fv


def search_local_points(frame_view, mp_view, th, nn_ratio):
    """C2 through the reference's own ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th) (ORBmatcher.cc:72-183, with
    Frame::GetFeaturesInArea / AssignFeaturesToGrid / PosInGrid of Frame.cc) -> (match_of_feature, nmatches), as
    pyoracle.search_local_points."""
    match = np.empty(max(frame_view.n, 1), np.int32)
    n = C.c_int(0)
    lib().ref_orb_search_local_points(C.byref(frame_view), C.byref(mp_view), C.c_float(th), C.c_float(nn_ratio), _p(match), C.byref(n))
    return match[:frame_view.n], n.value


def search_last_frame(cur_view, last_view, th, mono=False, check_orientation=True):
    """C3 through the reference's own ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono)
    (ORBmatcher.cc:1710-1879; cv::Mat float expressions through the stand-in's gemm) -> (match_of_feature, nmatches), as
    pyoracle.search_last_frame."""
    match = np.empty(max(cur_view.n, 1), np.int32)
    n = C.c_int(0)
    lib().ref_orb_search_last_frame(C.byref(cur_view), C.byref(last_view), C.c_float(th), C.c_int(int(mono)), C.c_int(int(check_orientation)),
                                    _p(match), C.byref(n))
    return match[:cur_view.n], n.value


def search_bow(a_view, b_view, mode, nn_ratio, check_orientation=True):
    """C6 / C7 through the reference's own ORBmatcher::SearchByBoW overloads (ORBmatcher.cc:247-407 mode 0: key frame A, frame B;
    :729-880 mode 1: key frames A, B) -> (match vector, nmatches), as pyoracle.search_bow."""
    sz = b_view.n if mode == 0 else a_view.n
    match = np.empty(max(sz, 1), np.int32)
    n = C.c_int(0)
    lib().ref_orb_search_bow(C.byref(a_view), C.byref(b_view), C.c_int(mode), C.c_float(nn_ratio), C.c_int(int(check_orientation)), _p(match),
                             C.byref(n))
    return match[:sz], n.value


def search_keyframe_points(cur_view, pt_view, log_sf, th, orb_dist, check_orientation=True):
    """C4 through the reference's own ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, sAlreadyFound, th, ORBdist)
    (ORBmatcher.cc:1891-2024, with MapPoint::PredictScale of MapPoint.cc; Ow comes from the frame's pose inside the function) ->
    (match_of_feature, nmatches), as pyoracle.search_keyframe_points."""
    match = np.empty(max(cur_view.n, 1), np.int32)
    n = C.c_int(0)
    lib().ref_orb_search_keyframe_points(C.byref(cur_view), C.byref(pt_view), C.c_float(log_sf), C.c_float(th), C.c_int(int(orb_dist)),
                                         C.c_int(int(check_orientation)), _p(match), C.byref(n))
    return match[:cur_view.n], n.value


def search_sim3_points(kf_view, pt_view, log_sf, th):
    """C5 through the reference's own ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, vpPoints, vpMatched, th)
    (ORBmatcher.cc:423-554, with KeyFrame::GetFeaturesInArea / IsInImage and MapPoint::PredictScale(dist, KeyFrame*)); the view's pose
    is Scw.  -> (match_of_feature, nmatches, tcw (12,) and ow (3,) after the function's own division by the scale, ORBmatcher.cc:435-439:
    what pyoracle.search_sim3_points / the C ABI take)."""
    match = np.empty(max(kf_view.n, 1), np.int32)
    n = C.c_int(0)
    tcw, ow = np.zeros(12, np.float32), np.zeros(3, np.float32)
    lib().ref_orb_search_sim3_points(C.byref(kf_view), C.byref(pt_view), C.c_float(log_sf), C.c_int(int(th)), _p(match), C.byref(n), _p(tcw), _p(ow))
    return match[:kf_view.n], n.value, tcw, ow


def line_search_by_projection(cur_view, line_view, local_map=False):
    """D3 through the reference's own LineMatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame)
    (LineMatcher.cpp:72-270, with LiangBarsky, UpdateKeyLineData, LineMatching, LineOverLap, ReprojectionError; Eigen and
    cv::LineIterator through stand-ins); local_map: D5, SearchByProjection(Frame& F, const vector<MapLine*>& vpMapLines) (:755-952)
    -> (match_of_line -> index of the last frame's / the map's line or -1, nmatches)."""
    match = np.empty(max(cur_view.n, 1), np.int32)
    n = C.c_int(0)
    lib().ref_line_search_by_projection(C.byref(cur_view), C.byref(line_view), C.c_int(int(local_map)), _p(match), C.byref(n))   # local_map: 0 = D3, 1 = D5, 2 = D4 (the key-frame overload, :527-753)
    return match[:cur_view.n], n.value


def frame_is_in_frustum_batch(tcw, ow, K, bounds, n_levels, log_sf, world_pos, normal, min_inv, max_inv, max_raw, cos_limit=0.5):
    """F4 through the reference's own Frame::IsInFrustum(MapPoint*, viewingCosLimit) (Frame.cc:345-401) for n frames x m map points ->
    (in_view, proj_x, proj_y, proj_xr, level, view_cos), as pyoracle.frame_is_in_frustum_batch (fields of points not in view are 0)."""
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
    for i in range(n):
        lib().ref_frame_is_in_frustum(_p(tcw[i]), _p(ow[i]), C.c_float(K["fx"]), C.c_float(K["fy"]), C.c_float(K["cx"]), C.c_float(K["cy"]),
                                      C.c_float(K["bf"]), _p(b), C.c_int(n_levels), C.c_float(log_sf), C.c_int(m), _p(wp), _p(no), _p(mi), _p(ma), _p(mr),
                                      C.c_float(cos_limit), _p(iv[i]), _p(px[i]), _p(py[i]), _p(pxr[i]), _p(lv[i]), _p(vc[i]))
    return iv, px, py, pxr, lv, vc


def frame_stereo_from_rgbd(depth, xy, x_un, bf):
    """F2 through the reference's own Frame::ComputeStereoFromRGBD (Frame.cc:1065-1117), one frame: depth (rows, cols) float32,
    xy = the raw key points, x_un = mvKeysUn[i].pt.x -> (mvDepth, mvuRight)."""
    depth = np.ascontiguousarray(depth, np.float32)
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    xu = np.ascontiguousarray(x_un, np.float32)
    d, ur = np.empty(max(len(xu), 1), np.float32), np.empty(max(len(xu), 1), np.float32)
    lib().ref_frame_stereo_from_rgbd(_p(depth), C.c_int(depth.shape[0]), C.c_int(depth.shape[1]), C.c_int(len(xu)), _p(xy), _p(xu), C.c_float(bf), _p(d), _p(ur))
    return d[:len(xu)], ur[:len(xu)]


def frame_unproject(xy_un, z, rwc, ow, K):
    """F3 through the reference's own Frame::UnprojectStereo (Frame.cc:1120-1134), one frame -> (world (n, 3), valid)."""
    xy = np.ascontiguousarray(xy_un, np.float32).reshape(-1, 2)
    z = np.ascontiguousarray(z, np.float32)
    r = np.ascontiguousarray(rwc, np.float32).reshape(9)
    o = np.ascontiguousarray(ow, np.float32).reshape(3)
    w, v = np.empty((max(len(z), 1), 3), np.float32), np.empty(max(len(z), 1), np.uint8)
    lib().ref_frame_unproject(C.c_int(len(z)), _p(xy), _p(z), _p(r), _p(o), C.c_float(K["fx"]), C.c_float(K["fy"]), C.c_float(K["cx"]), C.c_float(K["cy"]),
                              _p(w), _p(v))
    return w[:len(z)], v[:len(z)]


def search_for_initialization(f1, f2, prev_matched, window_size, nn_ratio=0.9, check_orientation=True):
    """ORBmatcher::SearchForInitialization (ORBmatcher.cc:573-717) through the reference's own function -> (vnMatches12, nmatches,
    vbPrevMatched after the call), as pyoracle.search_for_initialization."""
    pm = np.array(prev_matched, np.float32).reshape(-1, 2).copy()
    m = np.empty(max(f1.n, 1), np.int32)
    n = C.c_int(0)
    lib().ref_orb_search_for_initialization(C.byref(f1), C.byref(f2), _p(pm), C.c_int(int(window_size)), C.c_float(nn_ratio),
                                            C.c_int(int(check_orientation)), _p(m), C.byref(n))
    return m[:f1.n], n.value, pm


def distinctive_descriptors(desc, group_off):
    """MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:256-321) through the reference's own function, once per group of
    observations -> index of the chosen descriptor inside each group (-1 for an empty group)."""
    d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    off = np.ascontiguousarray(group_off, np.int32)
    best = np.empty(max(len(off) - 1, 1), np.int32)
    lib().ref_distinctive_descriptors(_p(d), _p(off), C.c_int(len(off) - 1), _p(best))
    return best[:len(off) - 1]


def fuse(kf_view, pt_view, ow, log_sf, inv_level_sigma2, th, variant=0):
    """ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, th) (ORBmatcher.cc:1107-1277) through the reference's own
    function -> (key-frame feature each map point was fused with or -1, nFused), as the first and third result of
    pyoracle.fuse_candidates(..., variant=0).  variant 1: the Sim3 overload Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) (:1290-1427), the
    view's pose being Scw; additionally returns the pose after the function's own division by the scale (tcw (12,), ow (3,))."""
    bi = np.empty(max(pt_view.n, 1), np.int32)
    n = C.c_int(0)
    o = np.ascontiguousarray(ow, np.float32)
    sg = np.ascontiguousarray(inv_level_sigma2, np.float32)
    tcw, ow2 = np.zeros(12, np.float32), np.zeros(3, np.float32)
    lib().ref_orb_fuse(C.byref(kf_view), C.byref(pt_view), _p(o), C.c_float(log_sf), _p(sg), C.c_float(th), C.c_int(int(variant)), _p(bi), C.byref(n),
                       _p(tcw), _p(ow2))
    return (bi[:pt_view.n], n.value) if variant == 0 else (bi[:pt_view.n], n.value, tcw, ow2)


def search_by_sim3(kf1, kf2, pts1, pts2, s12, R12, t12, log_sf1, log_sf2, th):
    """ORBmatcher::SearchBySim3 (ORBmatcher.cc:1441-1692) through the reference's own function -> (vpMatches12 as indices into key
    frame 2's features, nFound, [sR21 | t21] and [sR12 | t12] (12,) as the function forms them: what pyoracle.search_by_sim3 takes)."""
    r = np.ascontiguousarray(R12, np.float32).reshape(9)
    t = np.ascontiguousarray(t12, np.float32).reshape(3)
    m = np.empty(max(kf1.n, 1), np.int32)
    n = C.c_int(0)
    t21o, t12o = np.zeros(12, np.float32), np.zeros(12, np.float32)
    lib().ref_orb_search_by_sim3(C.byref(kf1), C.byref(kf2), C.byref(pts1), C.byref(pts2), C.c_float(s12), _p(r), _p(t), C.c_float(log_sf1),
                                 C.c_float(log_sf2), C.c_float(th), _p(m), C.byref(n), _p(t21o), _p(t12o))
    return m[:kf1.n], n.value, t21o, t12o
