"""ctypes loader of the in-tree native library (libplslam.so, sm_100a).

There is no fallback: if the library is missing or cannot be loaded, importing the product API fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libplslam.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
KL_DTYPE = np.dtype([("angle", "<f4"), ("class_id", "<i4"), ("octave", "<i4"), ("pt_x", "<f4"), ("pt_y", "<f4"),
                     ("response", "<f4"), ("size", "<f4"), ("sx", "<f4"), ("sy", "<f4"), ("ex", "<f4"), ("ey", "<f4"),
                     ("sx_oct", "<f4"), ("sy_oct", "<f4"), ("ex_oct", "<f4"), ("ey_oct", "<f4"), ("length", "<f4"),
                     ("num_pixels", "<i4")])
assert KP_DTYPE.itemsize == 28 and KL_DTYPE.itemsize == 68

PL_OK, PL_ERR_ARG, PL_ERR_EMPTY, PL_ERR_CUDA, PL_ERR_CAPACITY, PL_ERR_STATE = 0, -1, -2, -3, -4, -5


class PlError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"plslam error {code}: {msg}")
        self.code = code


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
                "There is no CPU or PyTorch fallback for this path.")
        _lib = C.CDLL(LIB_PATH)
        _lib.pl_last_error.restype = C.c_char_p
        _lib.pl_build_info.restype = C.c_char_p
        _lib.pl_orb_scale_factor.restype = C.c_float
        _lib.pl_orb_scale_factor.argtypes = [C.c_void_p]
        for name in ("pl_orb_stream", "pl_match_stream", "pl_line_stream"):
            if hasattr(_lib, name):
                getattr(_lib, name).restype = C.c_void_p
                getattr(_lib, name).argtypes = [C.c_void_p]
    return _lib


def check(rc: int):
    if rc != PL_OK:
        raise PlError(rc, lib().pl_last_error().decode("utf-8", "replace"))


def ptr(a):
    """void* of a numpy array or a raw integer device address."""
    if a is None:
        return C.c_void_p(0)
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------------------------------------------------------
# POD views of include/plslam_c.h (same layout for the native library and for the CPU oracle)
# ---------------------------------------------------------------------------------------------------------------
class FrameView(C.Structure):
    _fields_ = [("n", C.c_int), ("keys_un", C.c_void_p), ("desc", C.c_void_p), ("u_right", C.c_void_p), ("claimed", C.c_void_p),
                ("min_x", C.c_float), ("min_y", C.c_float), ("max_x", C.c_float), ("max_y", C.c_float),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("bf", C.c_float), ("b", C.c_float),
                ("tcw", C.c_float * 12), ("n_levels", C.c_int), ("scale_factors", C.c_void_p)]


class MapPointView(C.Structure):
    _fields_ = [("n", C.c_int), ("desc", C.c_void_p), ("track_in_view", C.c_void_p), ("proj_x", C.c_void_p), ("proj_y", C.c_void_p),
                ("proj_xr", C.c_void_p), ("scale_level", C.c_void_p), ("view_cos", C.c_void_p), ("has_observations", C.c_void_p)]


class LastFrameView(C.Structure):
    _fields_ = [("n", C.c_int), ("valid", C.c_void_p), ("world_pos", C.c_void_p), ("desc", C.c_void_p), ("octave", C.c_void_p),
                ("angle", C.c_void_p), ("has_observations", C.c_void_p), ("tcw", C.c_float * 12)]


def _addr(a):
    if a is None:
        return None
    # (__array_interface__ is the cheap way to the address, except for structured dtypes, whose interface spells out the fields)
    return a.ctypes.data if a.dtype.names else a.__array_interface__["data"][0]


def _c(a, dt):
    """np.ascontiguousarray(a, dt) without the call when `a` already is one (the glue builds hundreds of views per step)."""
    if type(a) is np.ndarray and a.dtype == dt and a.flags.c_contiguous:
        return a
    return np.ascontiguousarray(a, dt)


def _tcw12(tcw):
    t = _c(np.asarray(tcw).reshape(-1)[:12], np.float32)
    return (C.c_float * 12).from_buffer_copy(t)


def make_frame_view(keys_un, desc, u_right, claimed, bounds, K, tcw, scale_factors, keep):
    """keep: list that receives the arrays referenced by the view (lifetime)."""
    keys_un = _c(keys_un, KP_DTYPE)
    desc = _c(desc, np.uint8)
    u_right = _c(u_right, np.float32)
    claimed = None if claimed is None else _c(claimed, np.int32)
    sf = _c(scale_factors, np.float32)
    keep += [keys_un, desc, u_right, claimed, sf]
    v = FrameView()
    v.n = len(keys_un)
    v.keys_un, v.desc, v.u_right, v.claimed = _addr(keys_un), _addr(desc), _addr(u_right), _addr(claimed)
    v.min_x, v.min_y, v.max_x, v.max_y = [float(b) for b in bounds]
    v.fx, v.fy, v.cx, v.cy, v.bf = float(K["fx"]), float(K["fy"]), float(K["cx"]), float(K["cy"]), float(K["bf"])
    v.b = float(np.float32(K["bf"]) / np.float32(K["fx"]))
    v.tcw = _tcw12(tcw)
    v.n_levels = len(sf)
    v.scale_factors = _addr(sf)
    return v


def make_mappoint_view(desc, in_view, proj_x, proj_y, proj_xr, level, view_cos, has_obs, keep):
    desc = _c(desc, np.uint8)
    in_view = _c(in_view, np.uint8)
    px, py, pxr = (_c(a, np.float32) for a in (proj_x, proj_y, proj_xr))
    level = _c(level, np.int32)
    vc = _c(view_cos, np.float32)
    ho = None if has_obs is None else _c(has_obs, np.uint8)
    keep += [desc, in_view, px, py, pxr, level, vc, ho]
    v = MapPointView()
    v.n = len(in_view)
    v.desc, v.track_in_view, v.proj_x, v.proj_y, v.proj_xr = _addr(desc), _addr(in_view), _addr(px), _addr(py), _addr(pxr)
    v.scale_level, v.view_cos, v.has_observations = _addr(level), _addr(vc), _addr(ho)
    return v


class LocalMapView(C.Structure):
    """pl_localmap_view: one snapshot of the local map for pl_orb_search_local_map_batch (IsInFrustum + C2 on the device)."""
    _fields_ = [("n", C.c_int), ("world_pos", C.c_void_p), ("normal", C.c_void_p), ("desc", C.c_void_p), ("min_dist_inv", C.c_void_p),
                ("max_dist_inv", C.c_void_p), ("max_dist", C.c_void_p), ("has_observations", C.c_void_p)]


def make_localmap_view(world_pos, normal, desc, min_dist_inv, max_dist_inv, max_dist, has_obs, keep):
    wp = _c(world_pos, np.float32).reshape(-1, 3)
    no = _c(normal, np.float32).reshape(-1, 3)
    desc = _c(desc, np.uint8).reshape(-1, 32)
    mi, ma, mr = (_c(a, np.float32) for a in (min_dist_inv, max_dist_inv, max_dist))
    ho = None if has_obs is None else _c(has_obs, np.uint8)
    assert len(no) == len(wp) == len(desc) == len(mi) == len(ma) == len(mr)
    keep += [wp, no, desc, mi, ma, mr, ho]
    v = LocalMapView()
    v.n = len(wp)
    v.world_pos, v.normal, v.desc, v.min_dist_inv, v.max_dist_inv, v.max_dist = _addr(wp), _addr(no), _addr(desc), _addr(mi), _addr(ma), _addr(mr)
    v.has_observations = _addr(ho)
    return v


def make_lastframe_view(valid, world_pos, desc, octave, angle, has_obs, tcw, keep):
    valid = _c(valid, np.uint8)
    wp = _c(world_pos, np.float32)
    desc = _c(desc, np.uint8)
    octave = _c(octave, np.int32)
    angle = _c(angle, np.float32)
    ho = None if has_obs is None else _c(has_obs, np.uint8)
    keep += [valid, wp, desc, octave, angle, ho]
    v = LastFrameView()
    v.n = len(valid)
    v.valid, v.world_pos, v.desc, v.octave, v.angle, v.has_observations = _addr(valid), _addr(wp), _addr(desc), _addr(octave), _addr(angle), _addr(ho)
    v.tcw = _tcw12(tcw)
    return v


class MapLineView(C.Structure):
    _fields_ = [("n", C.c_int), ("start3d", C.c_void_p), ("end3d", C.c_void_p), ("kl", C.c_void_p), ("desc", C.c_void_p), ("valid", C.c_void_p)]


class LineFrameView(C.Structure):
    _fields_ = [("n", C.c_int), ("kl", C.c_void_p), ("desc", C.c_void_p), ("claimed", C.c_void_p), ("tcw", C.c_float * 12),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("min_x", C.c_float), ("min_y", C.c_float), ("max_x", C.c_float), ("max_y", C.c_float), ("cols", C.c_int), ("rows", C.c_int)]


def make_mapline_view(start3d, end3d, kl, desc, valid, keep):
    s3 = _c(start3d, np.float64).reshape(-1, 3)
    e3 = _c(end3d, np.float64).reshape(-1, 3)
    kl = _c(kl, KL_DTYPE)
    desc = _c(desc, np.uint8).reshape(-1, 32)
    valid = _c(valid, np.uint8)
    keep += [s3, e3, kl, desc, valid]
    v = MapLineView()
    v.n = len(kl)
    v.start3d, v.end3d, v.kl, v.desc, v.valid = _addr(s3), _addr(e3), _addr(kl), _addr(desc), _addr(valid)
    return v


def make_lineframe_view(kl, desc, claimed, tcw, K, bounds, img_size, keep):
    kl = _c(kl, KL_DTYPE)
    desc = _c(desc, np.uint8).reshape(-1, 32)
    claimed = None if claimed is None else _c(claimed, np.uint8)
    keep += [kl, desc, claimed]
    v = LineFrameView()
    v.n = len(kl)
    v.kl, v.desc, v.claimed = _addr(kl), _addr(desc), _addr(claimed)
    v.tcw = _tcw12(tcw)
    v.fx, v.fy, v.cx, v.cy = float(K["fx"]), float(K["fy"]), float(K["cx"]), float(K["cy"])
    v.min_x, v.min_y, v.max_x, v.max_y = [float(b) for b in bounds]
    v.cols, v.rows = int(img_size[0]), int(img_size[1])
    return v


class PosePointView(C.Structure):
    """pl_posepoint_view: map points offered to the pose-based projection searches (C4 / C5)."""
    _fields_ = [("n", C.c_int), ("valid", C.c_void_p), ("world_pos", C.c_void_p), ("desc", C.c_void_p), ("min_dist_inv", C.c_void_p),
                ("max_dist_inv", C.c_void_p), ("max_dist", C.c_void_p), ("angle", C.c_void_p), ("normal", C.c_void_p)]


def make_posepoint_view(valid, world_pos, desc, min_dist_inv, max_dist_inv, max_dist, angle, normal, keep):
    valid = _c(valid, np.uint8)
    wp = _c(world_pos, np.float32).reshape(-1, 3)
    desc = _c(desc, np.uint8).reshape(-1, 32)
    mi = _c(min_dist_inv, np.float32)
    ma = _c(max_dist_inv, np.float32)
    mr = _c(max_dist, np.float32)
    an = None if angle is None else _c(angle, np.float32)
    no = None if normal is None else _c(normal, np.float32).reshape(-1, 3)
    keep += [valid, wp, desc, mi, ma, mr, an, no]
    v = PosePointView()
    v.n = len(valid)
    v.valid, v.world_pos, v.desc, v.min_dist_inv, v.max_dist_inv, v.max_dist = _addr(valid), _addr(wp), _addr(desc), _addr(mi), _addr(ma), _addr(mr)
    v.angle, v.normal = _addr(an), _addr(no)
    return v


class BowView(C.Structure):
    """pl_bow_view: one side of ORBmatcher::SearchByBoW (features + flattened DBoW2::FeatureVector)."""
    _fields_ = [("n", C.c_int), ("angle", C.c_void_p), ("desc", C.c_void_p), ("valid", C.c_void_p), ("n_nodes", C.c_int),
                ("node_id", C.c_void_p), ("node_off", C.c_void_p), ("feat_idx", C.c_void_p)]


def make_bow_view(angle, desc, valid, feat_vec, keep):
    """feat_vec: dict node id -> list of feature indices (DBoW2::FeatureVector); flattened in key order."""
    angle = _c(angle, np.float32)
    desc = _c(desc, np.uint8).reshape(-1, 32)
    valid = None if valid is None else _c(valid, np.uint8)
    ids = sorted(feat_vec.keys())
    node_id = np.asarray(ids, np.uint32)
    off = np.zeros(len(ids) + 1, np.int32)
    feats = []
    for k, nid in enumerate(ids):
        feats.extend(int(x) for x in feat_vec[nid])
        off[k + 1] = len(feats)
    feat_idx = np.asarray(feats, np.uint32)
    keep += [angle, desc, valid, node_id, off, feat_idx]
    v = BowView()
    v.n = len(angle)
    v.angle, v.desc, v.valid = _addr(angle), _addr(desc), _addr(valid)
    v.n_nodes = len(ids)
    v.node_id, v.node_off, v.feat_idx = _addr(node_id), _addr(off), _addr(feat_idx)
    return v


class TriangView(C.Structure):
    """pl_triang_view: one key frame of ORBmatcher::SearchForTriangulation (bow view + undistorted keys + mvuRight)."""
    _fields_ = [("bow", BowView), ("keys_un", C.c_void_p), ("u_right", C.c_void_p)]


def make_triang_view(keys_un, desc, u_right, no_mappoint, feat_vec, keep):
    """no_mappoint[i] = 1 when the key frame holds no map point for feature i (the features offered for triangulation)."""
    keys_un = _c(keys_un, KP_DTYPE)
    ur = _c(u_right, np.float32)
    ang = np.ascontiguousarray(keys_un["angle"], np.float32)
    keep += [keys_un, ur]
    v = TriangView()
    v.bow = make_bow_view(ang, desc, no_mappoint, feat_vec, keep)
    v.keys_un, v.u_right = _addr(keys_un), _addr(ur)
    return v
