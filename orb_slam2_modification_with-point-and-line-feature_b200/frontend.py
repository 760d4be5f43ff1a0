"""Host-side glue that stands in for the reference's callers of the hot path (Frame / Tracking), in numpy.

The reference's `Frame` constructor (src/Frame.cc:135-205) runs the two extractors, looks depth up for every
keypoint (ComputeStereoFromRGBD, :1065-1117) and `Tracking` (src/Tracking.cc) feeds the matchers with the last frame
(TrackWithMotionModel, :1212-1290) and the local map (SearchLocalPoints / SearchLocalLines, :1746-1865).  Those callers
are OUT of the hot-path scope; this module reproduces just enough of them to drive the extractors and matchers on a
synthetic RGB-D sequence.  The SAME glue drives both arms — the CUDA library through the C
ABI and the CPU oracle — so both see byte-identical inputs.

Backends expose: extract_orb(frames) -> [(kps, desc)], extract_lines(frames) -> [(kls, desc, coeffs)],
search_last_frame(cur_view, last_view, th), search_local_points(frame_view, mp_view, th, nn_ratio) and their *_batch
forms over lists of views, line_search_batch(line_frame_views, map_line_views).
"""
from __future__ import annotations

import time

import numpy as np

from . import _native as N
from .synth import TUM1

f32 = np.float32


class FeatureList(list):
    """A list of per-frame feature tuples that also carries the dense arrays they are views of — `dense` = (array (n, cap) of the
    structured key points / key lines, counts (n,)) — so that the sequence-wide parts of the glue gather from the dense array once
    instead of concatenating n small views."""
    dense = None


class FrameLite:
    """The subset of ORB_SLAM2::Frame the matchers read."""

    def __init__(self, kps, desc, kls, ldesc, depth, Tcw, K, scale_factors, stereo=None):
        self.kps, self.desc, self.kls, self.ldesc = kps, desc, kls, ldesc
        self.Tcw = np.asarray(Tcw, f32)
        self.K = K
        self.sf = np.asarray(scale_factors, f32)
        h, w = depth.shape
        self.bounds = (0.0, 0.0, float(w), float(h))  # ComputeImageBounds without distortion (Frame.cc:852-887)
        self.size = (w, h)
        self._world = None
        if stereo is not None:  # (mvDepth, mvuRight) already computed for the whole sequence (build_batch)
            self.depth, self.u_right = stereo
        else:
            # ComputeStereoFromRGBD (Frame.cc:1065-1117): depth at the (truncated) keypoint position
            u = kps["x"].astype(np.int64)
            v = kps["y"].astype(np.int64)
            d = depth[np.clip(v, 0, h - 1), np.clip(u, 0, w - 1)].astype(f32)
            ok = d > 0
            self.depth = np.where(ok, d, f32(-1)).astype(f32)
            with np.errstate(divide="ignore", invalid="ignore"):
                self.u_right = np.where(ok, kps["x"] - f32(K["bf"]) / d, f32(-1)).astype(f32)
        Rcw, tcw = self.Tcw[:3, :3], self.Tcw[:3, 3]
        self.Rwc = Rcw.T.copy()
        self.Ow = (-self.Rwc @ tcw).astype(f32)
        self._depth_img = depth
        self.set_lines(kls, ldesc)

    @staticmethod
    def build_batch(orb, depth, Tcw, K, scale_factors, backend=None):
        """The Frame constructors of a whole sequence: ComputeStereoFromRGBD as one gather over all key points and UnprojectStereo
        of every frame in one pass — through the backend's batched F3 call (pl_frame_unproject_batch / its oracle twin) when a
        backend is given, else in numpy (same formula; used while the device is saturated by the line extractor)."""
        n = len(orb)
        counts = np.array([len(o[0]) for o in orb], np.int64)
        off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        if n == 0 or not isinstance(depth, np.ndarray) or depth.ndim != 3:
            return [FrameLite(orb[t][0], orb[t][1], None, None, depth[t], Tcw[t], K, scale_factors) for t in range(n)]
        dense = getattr(orb, "dense", None)
        if dense is not None:
            mask = np.arange(dense[0].shape[1])[None, :] < counts[:, None]
            x, y = dense[0]["x"][mask], dense[0]["y"][mask]
        else:
            x = np.concatenate([o[0]["x"] for o in orb])
            y = np.concatenate([o[0]["y"] for o in orb])
        fidx = np.repeat(np.arange(n), counts)
        h, w = depth.shape[1:]
        d = depth[fidx, np.clip(y.astype(np.int64), 0, h - 1), np.clip(x.astype(np.int64), 0, w - 1)].astype(f32)
        ok = d > 0
        dd = np.where(ok, d, f32(-1)).astype(f32)
        with np.errstate(divide="ignore", invalid="ignore"):
            ur = np.where(ok, x - f32(K["bf"]) / d, f32(-1)).astype(f32)
        frames = [FrameLite(orb[t][0], orb[t][1], None, None, depth[t], Tcw[t], K, scale_factors, stereo=(dd[off[t]:off[t + 1]], ur[off[t]:off[t + 1]]))
                  for t in range(n)]
        rwc = np.stack([F.Rwc for F in frames])
        ow = np.stack([F.Ow for F in frames])
        if backend is not None and hasattr(backend, "unproject_batch"):
            world, _ = backend.unproject_batch(off, np.stack([x, y], 1), dd, rwc, ow, K)
        else:
            pc = np.stack([(x - f32(K["cx"])) * dd * f32(1.0 / K["fx"]), (y - f32(K["cy"])) * dd * f32(1.0 / K["fy"]), dd], 1).astype(f32)
            world = (np.einsum("nij,nj->ni", rwc[fidx], pc) + ow[fidx]).astype(f32)
        for t, F in enumerate(frames):
            F._world = world[off[t]:off[t + 1]]
        return frames

    def set_lines(self, kls, ldesc):
        """Attaches the line features (they may arrive later than the points: the two extractors run concurrently)."""
        self.kls, self.ldesc = kls, ldesc
        self._ul = None
        depth = self._depth_img
        h, w = depth.shape
        # line endpoints: depth at the rounded endpoint
        if kls is not None and len(kls):
            sx = np.clip(np.rint(kls["sx"]).astype(np.int64), 0, w - 1)
            sy = np.clip(np.rint(kls["sy"]).astype(np.int64), 0, h - 1)
            ex = np.clip(np.rint(kls["ex"]).astype(np.int64), 0, w - 1)
            ey = np.clip(np.rint(kls["ey"]).astype(np.int64), 0, h - 1)
            self.ds, self.de = depth[sy, sx].astype(np.float64), depth[ey, ex].astype(np.float64)
        else:
            self.ds = self.de = np.zeros(0)

    # Frame::UnprojectStereo (Frame.cc:1120-1134), vectorised
    def unproject_points(self):
        if self._world is not None:
            return self._world
        z = self.depth
        x = (self.kps["x"] - f32(self.K["cx"])) * z * f32(1.0 / self.K["fx"])
        y = (self.kps["y"] - f32(self.K["cy"])) * z * f32(1.0 / self.K["fy"])
        pc = np.stack([x, y, z], 1).astype(f32)
        return (pc @ self.Rwc.T + self.Ow).astype(f32)

    def unproject_lines(self):
        if getattr(self, "_ul", None) is not None:
            return self._ul
        K = self.K
        Rwc, Ow = self.Rwc.astype(np.float64), self.Ow.astype(np.float64)

        def lift(px, py, z):
            pc = np.stack([(px - K["cx"]) * z / K["fx"], (py - K["cy"]) * z / K["fy"], z], 1)
            return pc @ Rwc.T + Ow

        s3 = lift(self.kls["sx"].astype(np.float64), self.kls["sy"].astype(np.float64), self.ds)
        e3 = lift(self.kls["ex"].astype(np.float64), self.kls["ey"].astype(np.float64), self.de)
        self._ul = (s3, e3, (self.ds > 0) & (self.de > 0))
        return self._ul

    @staticmethod
    def prepare_lines_batch(frames, depth):
        """The part of attach_lines_batch that does not need the line features (pose stacks, the flat depth view): the caller runs
        it while the line extractor is still working."""
        if not len(frames) or not isinstance(depth, np.ndarray) or depth.ndim != 3:
            return None
        n = len(frames)
        h, w = depth.shape[1:]
        Rwc = np.stack([F.Rwc for F in frames]).astype(np.float64)
        return dict(n=n, RT=np.ascontiguousarray(Rwc.transpose(0, 2, 1)), Ow=np.stack([F.Ow for F in frames]).astype(np.float64)[:, None, :],
                    tcw=np.stack([F.Tcw[:3].reshape(-1) for F in frames]).astype(f32),
                    base=(np.arange(n, dtype=np.int64) * (h * w))[:, None], flat=depth.reshape(-1))

    @staticmethod
    def attach_lines_batch(frames, lines, depth, prep=None):
        """set_lines + unproject_lines of a whole sequence in one vectorised pass (depth: the (n, h, w) array the frames were
        built from).  Same arithmetic as the per-frame methods."""
        n = len(frames)
        dense = getattr(lines, "dense", None)
        if dense is not None and len(dense) > 1 and len(dense[1]) == n:
            counts = np.asarray(dense[1]).astype(np.int64)  # the extractor's own per-frame counts
        else:
            counts = np.array([len(l[0]) for l in lines], np.int64)
        if counts.sum() == 0 or not isinstance(depth, np.ndarray) or depth.ndim != 3:
            for F, (kls, ldesc, _) in zip(frames, lines):
                F.set_lines(kls, ldesc)
            return
        h, w = depth.shape[1:]
        K = frames[0].K
        if prep is None or prep["n"] != n:
            prep = FrameLite.prepare_lines_batch(frames, depth)
        if dense is not None and dense[0].ndim == 2 and len(dense[0]) == n:
            # the extractor's own (n, max_lines) output: every row at once (rows beyond a frame's count are never looked at);
            # the batched matmul is the per-frame `pc @ Rwc.T` of unproject_lines, row for row
            kl = dense[0]
            m = kl.shape[1]
            # start and end points go through the arithmetic together: xy[0] = (sx, sy), xy[1] = (ex, ey) of every line
            xy = np.empty((2, n, m, 2), np.float64)
            if kl.dtype == N.KL_DTYPE and kl.flags.c_contiguous:
                raw = kl.view(np.float32).reshape(n, m, N.KL_DTYPE.itemsize // 4)   # sx, sy, ex, ey are words 7..10 of a KeyLine
                xy[0], xy[1] = raw[:, :, 7:9], raw[:, :, 9:11]
            else:
                xy[0, :, :, 0], xy[0, :, :, 1], xy[1, :, :, 0], xy[1, :, :, 1] = kl["sx"], kl["sy"], kl["ex"], kl["ey"]
            px, py = xy[..., 0], xy[..., 1]
            with np.errstate(invalid="ignore"):
                ix = np.minimum(np.maximum(np.rint(xy).astype(np.int64), 0), np.array([w - 1, h - 1], np.int64))   # np.clip per axis
            base, flat, RT, Ow = prep["base"], prep["flat"], prep["RT"], prep["Ow"]
            z = flat[base[None] + ix[..., 1] * w + ix[..., 0]].astype(np.float64)   # depth at the rounded end points
            with np.errstate(invalid="ignore", over="ignore"):
                pc = np.empty((2, n, m, 3), np.float64)
                pc[..., 0] = (px - K["cx"]) * z / K["fx"]
                pc[..., 1] = (py - K["cy"]) * z / K["fy"]
                pc[..., 2] = z
                P3 = np.matmul(pc, RT[None]) + Ow[None]
                s3, e3, ds, de = P3[0], P3[1], z[0], z[1]
                ok = (ds > 0) & (de > 0)
            for t, (F, c) in enumerate(zip(frames, counts.tolist())):
                F.kls, F.ldesc = lines[t][0], lines[t][1]
                F.ds, F.de = ds[t, :c], de[t, :c]
                F._ul = (s3[t, :c], e3[t, :c], ok[t, :c])
            ld = dense[2] if len(dense) > 2 else None
            if (ld is not None and kl.flags.c_contiguous and ld.flags.c_contiguous and ld.dtype == np.uint8 and ld.shape[:2] == kl.shape
                    and kl.dtype == N.KL_DTYPE):
                # what the line searches read, as whole-sequence arrays: their views are then built for all frames at once
                frames[0]._seq = dict(kl=kl, ld=ld, s3=np.ascontiguousarray(s3), e3=np.ascontiguousarray(e3), ok=ok.astype(np.uint8), counts=counts,
                                      tcw=prep["tcw"])
            return
        for F, (kls, ldesc, _) in zip(frames, lines):  # features handed over as a list of per-frame arrays
            F.set_lines(kls, ldesc)

    def view(self, claimed, keep):
        return N.make_frame_view(self.kps, self.desc, self.u_right, claimed, self.bounds, self.K, self.Tcw[:3].reshape(-1), self.sf, keep)


class LocalMap:
    """Map points / lines created from depth on keyframes (Tracking.cc:633-692), with the fields IsInFrustum needs."""

    def __init__(self, max_points=3000, max_lines=300):
        self.max_points, self.max_lines = max_points, max_lines
        self.pos = np.zeros((0, 3), f32)
        self.desc = np.zeros((0, 32), np.uint8)
        self.normal = np.zeros((0, 3), f32)
        self.max_d = np.zeros(0, f32)
        self.min_d = np.zeros(0, f32)
        self.ls = np.zeros((0, 3))
        self.le = np.zeros((0, 3))
        self.lkl = np.zeros(0, N.KL_DTYPE)
        self.ldesc = np.zeros((0, 32), np.uint8)

    def add_keyframe(self, F: FrameLite):
        self.add_keyframe_points(F)
        self.add_keyframe_lines(F)

    def add_keyframe_points(self, F: FrameLite):
        ok = F.depth > 0
        P = F.unproject_points()[ok]
        PO = P - F.Ow
        dist = np.linalg.norm(PO, axis=1).astype(f32)
        octv = F.kps["octave"][ok]
        max_d = dist * F.sf[octv]                      # MapPoint::MapPoint (MapPoint.cc:57-66)
        min_d = max_d / F.sf[len(F.sf) - 1]
        self.pos = np.concatenate([self.pos, P])[-self.max_points:]
        self.desc = np.concatenate([self.desc, F.desc[ok]])[-self.max_points:]
        self.normal = np.concatenate([self.normal, (PO / dist[:, None]).astype(f32)])[-self.max_points:]
        self.max_d = np.concatenate([self.max_d, max_d])[-self.max_points:]
        self.min_d = np.concatenate([self.min_d, min_d])[-self.max_points:]

    def add_keyframe_lines(self, F: FrameLite):
        if len(F.kls):
            s3, e3, okl = F.unproject_lines()
            self.ls = np.concatenate([self.ls, s3[okl]])[-self.max_lines:]
            self.le = np.concatenate([self.le, e3[okl]])[-self.max_lines:]
            self.lkl = np.concatenate([self.lkl, F.kls[okl]])[-self.max_lines:]
            self.ldesc = np.concatenate([self.ldesc, F.ldesc[okl]])[-self.max_lines:]

    @staticmethod
    def line_snapshots_dense(seq, n, kf_every, max_lines=300):
        """lmaps[t] = (ls, le, lkl, ldesc) of the local line map frame t sees, for a whole sequence at once: what calling
        add_keyframe_lines on every kf_every-th frame leaves behind.  The map after a key frame is the trailing max_lines rows of
        everything the key frames so far contributed, so every snapshot is a slice of ONE gathered array (no per-key-frame
        concatenation of structured arrays)."""
        kf = np.arange(0, n, kf_every)
        sel = (np.arange(seq["kl"].shape[1])[None, :] < seq["counts"][kf][:, None]) & (seq["ok"][kf] != 0)
        ls, le, lkl, ld = seq["s3"][kf][sel], seq["e3"][kf][sel], seq["kl"][kf][sel], seq["ld"][kf][sel]
        hi = np.concatenate([[0], np.cumsum(sel.sum(1))]).astype(np.int64)   # snapshot k = the map after k key frames
        lo = np.maximum(hi - max_lines, 0)
        snaps = [(ls[a:b], le[a:b], lkl[a:b], ld[a:b]) for a, b in zip(lo.tolist(), hi.tolist())]
        snap_of = np.zeros(n, np.int64)
        snap_of[1:] = (np.arange(1, n) - 1) // kf_every + 1
        where = dict(ls=ls, le=le, lkl=lkl, ld=ld, lo=lo, hi=hi, snap_of=snap_of, ones=np.ones(max(max_lines, 1), np.uint8))
        return [snaps[k] for k in snap_of.tolist()], where

    # Frame::IsInFrustum (Frame.cc:345-401) + MapPoint::PredictScale (MapPoint.cc:416-431), vectorised float32 over the
    # frames that see the same snapshot of the map
    def frustum_group(self, Fs, cos_limit=0.5, backend=None):
        K = Fs[0].K
        sf = Fs[0].sf
        if backend is not None and hasattr(backend, "is_in_frustum_batch"):
            # F4: pl_frame_is_in_frustum_batch / its oracle twin — Frame::IsInFrustum with the reference's own rounding
            tcw = np.stack([F.Tcw[:3].reshape(-1) for F in Fs])
            ow = np.stack([F.Ow for F in Fs])
            log_sf = float(f32(np.log(f32(sf[1] / sf[0]))))
            iv, u, v, xr, lvl, vc = backend.is_in_frustum_batch(tcw, ow, K, Fs[0].bounds, len(sf), log_sf, self.pos, self.normal, self.min_d, self.max_d,
                                                                 self.max_d, cos_limit)
            return [(iv[i], u[i], v[i], xr[i], lvl[i], vc[i]) for i in range(len(Fs))]
        R = np.stack([F.Tcw[:3, :3] for F in Fs]).astype(f32)
        t = np.stack([F.Tcw[:3, 3] for F in Fs]).astype(f32)
        Ow = np.stack([F.Ow for F in Fs]).astype(f32)
        Pc = (np.matmul(self.pos[None], R.transpose(0, 2, 1)) + t[:, None, :]).astype(f32)
        z = Pc[..., 2]
        with np.errstate(divide="ignore", invalid="ignore"):
            invz = (f32(1.0) / z).astype(f32)
            u = (f32(K["fx"]) * Pc[..., 0] * invz + f32(K["cx"])).astype(f32)
            v = (f32(K["fy"]) * Pc[..., 1] * invz + f32(K["cy"])).astype(f32)
            PO = self.pos[None, :, :] - Ow[:, None, :]
            dist = np.sqrt(np.einsum("kmi,kmi->km", PO, PO)).astype(f32)
            view_cos = (np.einsum("kmi,mi->km", PO, self.normal) / dist).astype(f32)
            ratio = self.max_d[None] / dist
            lvl = np.ceil(np.log(ratio) / np.log(sf[1] / sf[0]))
        b = Fs[0].bounds
        ok = (z > 0) & (u >= b[0]) & (u <= b[2]) & (v >= b[1]) & (v <= b[3])
        ok &= (dist >= self.min_d[None]) & (dist <= self.max_d[None]) & (view_cos >= cos_limit)
        lvl = np.clip(np.nan_to_num(lvl), 0, len(sf) - 1).astype(np.int32)
        xr = (u - f32(K["bf"]) * invz).astype(f32)
        ok8, u, v, xr, vc = ok.astype(np.uint8), np.nan_to_num(u), np.nan_to_num(v), np.nan_to_num(xr), np.nan_to_num(view_cos)
        return [(ok8[i], u[i], v[i], xr[i], lvl[i], vc[i]) for i in range(len(Fs))]

    def frustum(self, F: FrameLite, cos_limit=0.5):
        return self.frustum_group([F], cos_limit)[0]


class TrackingFrontEnd:
    """Matching schedule of config 2: per frame C3 + D3 against the previous frame and C2 + D5 against the local map.

    The pose prior of every frame is known up front (ground truth + noise; no optimiser in the loop), so the caller
    state of all frames can be prepared first and each kind of search is issued as ONE batched call over the sequence
    (`batch=True`, the offline / throughput mode) or frame by frame (`batch=False`, the streaming mode).  Both modes
    run the same searches on the same inputs and return the same summary."""

    def __init__(self, backend, K=TUM1, keyframe_every=10, device_glue=False, fused_local_map=True):
        """device_glue: run the Frame glue (UnprojectStereo, IsInFrustum) through the backend's batched F-row calls instead of
        numpy.  Off by default for the throughput runs: while the ordered LSD stage owns every SM, work queued on the device
        waits for it, whereas the host cores are idle — the numpy path overlaps the line extraction, the device path queues
        behind it (tools/prof_e2e.py shows both timelines)."""
        self.b = backend
        self.K = K
        self.kf_every = keyframe_every
        self.device_glue = device_glue
        self.fused_local_map = fused_local_map
        self.keepalive = []  # arrays referenced by the views handed to the backend (views hold raw addresses)

    _chk_w = {}
    _obs = {}

    @staticmethod
    def _has_obs(n):
        """two of three key points of the last frame carry a map point with observations (the same pattern for every frame)"""
        a = TrackingFrontEnd._obs.get(n)
        if a is None:
            a = TrackingFrontEnd._obs[n] = (np.arange(n) % 3 != 0).astype(np.uint8)
        return a

    @staticmethod
    def _chk(m):
        w = TrackingFrontEnd._chk_w.get(len(m))
        if w is None:
            w = TrackingFrontEnd._chk_w[len(m)] = np.arange(1, len(m) + 1, dtype=np.int64)
        return int(np.dot(m.astype(np.int64), w) + (len(m) * (len(m) + 1)) // 2)   # sum((m + 1) * (i + 1))

    def _dense_lineframe_views(self, seq, tt, F0, keep):
        """pl_lineframe_view of frames tt, built as one structured array (same fields as N.make_lineframe_view fills)."""
        a = np.zeros(len(tt), np.dtype(N.LineFrameView))
        a["n"] = seq["counts"][tt]
        a["kl"] = seq["kl"].ctypes.data + tt * seq["kl"].strides[0]
        a["desc"] = seq["ld"].ctypes.data + tt * seq["ld"].strides[0]
        a["tcw"] = seq["tcw"][tt]
        for k in ("fx", "fy", "cx", "cy"):
            a[k] = float(self.K[k])
        a["min_x"], a["min_y"], a["max_x"], a["max_y"] = [float(b) for b in F0.bounds]
        a["cols"], a["rows"] = int(F0.size[0]), int(F0.size[1])
        keep += [a, seq]
        return (N.LineFrameView * len(tt)).from_buffer(a)

    @staticmethod
    def _dense_mapline_views(seq, tt, keep):
        """pl_mapline_view of the lines of frames tt lifted to 3-D (the `last` side of D3)."""
        a = np.zeros(len(tt), np.dtype(N.MapLineView))
        a["n"] = seq["counts"][tt]
        a["start3d"] = seq["s3"].ctypes.data + tt * seq["s3"].strides[0]
        a["end3d"] = seq["e3"].ctypes.data + tt * seq["e3"].strides[0]
        a["kl"] = seq["kl"].ctypes.data + tt * seq["kl"].strides[0]
        a["desc"] = seq["ld"].ctypes.data + tt * seq["ld"].strides[0]
        a["valid"] = seq["ok"].ctypes.data + tt * seq["ok"].strides[0]
        keep += [a, seq]
        return (N.MapLineView * len(tt)).from_buffer(a)

    @staticmethod
    def _dense_snapshot_views(where, tt, keep):
        """pl_mapline_view of the local line map frames tt see (the `map` side of D5): rows [lo, hi) of the gathered arrays of
        LocalMap.line_snapshots_dense; frames between two key frames point at the same rows."""
        k = where["snap_of"][tt]
        lo, hi = where["lo"][k], where["hi"][k]
        a = np.zeros(len(tt), np.dtype(N.MapLineView))
        a["n"] = hi - lo
        a["start3d"] = where["ls"].ctypes.data + lo * where["ls"].strides[0]
        a["end3d"] = where["le"].ctypes.data + lo * where["le"].strides[0]
        a["kl"] = where["lkl"].ctypes.data + lo * where["lkl"].strides[0]
        a["desc"] = where["ld"].ctypes.data + lo * where["ld"].strides[0]
        a["valid"] = where["ones"].ctypes.data
        keep += [a, where]
        return (N.MapLineView * len(tt)).from_buffer(a)

    @staticmethod
    def _line_results(summary, ts, res, tag):
        """(match_of_line, nmatches, used_relaxed, n_projected) of every search into the summary; a result that carries its dense
        arrays is summarised for all frames at once (rows are padded with -1, which adds nothing to sum((m + 1) * (i + 1)))."""
        dense = getattr(res, "dense", None)
        if dense is not None:
            out, cnt, rel, npj = dense
            chk = ((out.astype(np.int64) + 1) * np.arange(1, out.shape[1] + 1, dtype=np.int64)).sum(1) if len(out) else out
            for t, c, nl, r, p_ in zip(ts, chk.tolist(), cnt.tolist(), rel.tolist(), npj.tolist()):
                summary[t].update({tag + "_proj": p_, tag + "_matches": nl, tag + "_relaxed": r, tag + "_sum": c})
            return
        for t, (ml, nl, rel, npj) in zip(ts, res):
            summary[t].update({tag + "_proj": npj, tag + "_matches": nl, tag + "_relaxed": rel, tag + "_sum": TrackingFrontEnd._chk(ml)})

    def run(self, gray, depth, Tcw, scale_factors, features=None, prior_noise=True, batch=True):
        """features = (orb, lines); `lines` may be a concurrent.futures.Future: the point side (Frame-lite, C3, C2) does not
        need the lines, so it proceeds while the line extractor is still running — the reference runs its two extractors
        in two threads for the same reason (Frame.cc:152-155)."""
        n = len(gray)
        trace = getattr(self, "trace", None)   # optional list of (label, perf_counter) marks: where an end-to-end step goes
        mark = (lambda s_: trace.append((s_, time.perf_counter()))) if trace is not None else (lambda s_: None)
        mark("start")
        if features is None:
            orb = self.b.extract_orb(gray)
            lines = self.b.extract_lines(gray)
        else:
            orb, lines = features
        rng = np.random.Generator(np.random.PCG64(424242))
        keep = self.keepalive = []
        # ---- point side of the caller state (Frame-lite, local-map snapshots) ----
        maps = []
        lm = LocalMap()
        Tn = []
        for t in range(n):
            T = np.array(Tcw[t], np.float64)
            if prior_noise:  # pose prior = ground truth + small noise (stand-in for the motion model)
                T[:3, 3] += rng.normal(0, 0.002, 3)
            Tn.append(T.astype(f32))
        frames = FrameLite.build_batch(orb, depth, Tn, self.K, scale_factors, self.b if self.device_glue else None)
        for t in range(n):
            maps.append((lm.pos, lm.desc, lm.normal, lm.max_d, lm.min_d))
            if t % self.kf_every == 0:
                lm.add_keyframe_points(frames[t])
        summary = [dict(frame=t, n_kp=len(frames[t].kps)) for t in range(n)]
        mark("frames built")
        # ---- C3: ORBmatcher(0.9).SearchByProjection(Cur, Last, th=15) (Tracking.cc:1244) ----
        c3_t = list(range(1, n))
        cvs = [frames[t].view(None, keep) for t in c3_t]
        cv_of = dict(zip(c3_t, cvs))
        lvs = []
        for t in c3_t:
            last = frames[t - 1]
            lvs.append(N.make_lastframe_view(last.depth > 0, last.unproject_points(), last.desc, last.kps["octave"], last.kps["angle"],
                                             self._has_obs(len(last.kps)), last.Tcw[:3].reshape(-1), keep))
        # C2 inputs that do not depend on the C3 result (Frame::IsInFrustum of the local map): host work that overlaps the line
        # extraction still running on the device — and, when the point searches have a thread of their own, the C3 search
        c2_t = [t for t in range(n) if len(maps[t][0])]
        fr = []
        mark("C3 views built")

        def frustum_of_local_map():
            tmp = LocalMap()
            k = 0
            while k < len(c2_t):  # consecutive frames that see the same map snapshot are projected together
                k1 = k
                while k1 < len(c2_t) and maps[c2_t[k1]][0] is maps[c2_t[k]][0]:
                    k1 += 1
                tmp.pos, tmp.desc, tmp.normal, tmp.max_d, tmp.min_d = maps[c2_t[k]][:5]
                fr.extend((tmp.desc,) + r for r in tmp.frustum_group([frames[t] for t in c2_t[k:k1]], backend=self.b if self.device_glue else None))
                k = k1
            mark("frustum done")

        import threading
        fr_ready = threading.Event()
        # Tracking::SearchLocalPoints as ONE call (IsInFrustum + C2, the projections staying on the device) when the backend has it
        fused = batch and self.fused_local_map and hasattr(self.b, "search_local_map_batch")

        def point_searches():
            r3 = self.b.search_last_frame_batch(cvs, lvs, 15.0) if batch else [self.b.search_last_frame(c, l, 15.0) for c, l in zip(cvs, lvs)]
            claimed = [None] * n
            for t, (m3, n3) in zip(c3_t, r3):
                summary[t].update(c3_matches=n3, c3_sum=self._chk(m3))
                claimed[t] = (m3 >= 0).astype(np.int32)
            mark("  [points] C3 done")
            if fused:
                fvs, snaps, snap_of, mof = [], [], {}, []
                for t in c2_t:
                    if t in cv_of:
                        fv = type(cv_of[t]).from_buffer_copy(cv_of[t])
                        keep.append(claimed[t])
                        fv.claimed = claimed[t].__array_interface__["data"][0]
                        fvs.append(fv)
                    else:
                        fvs.append(frames[t].view(claimed[t], keep))
                    pos, desc, normal, max_d, min_d = maps[t][:5]
                    if id(pos) not in snap_of:
                        snap_of[id(pos)] = len(snaps)
                        snaps.append(N.make_localmap_view(pos, normal, desc, min_d, max_d, max_d, None, keep))
                    mof.append(snap_of[id(pos)])
                sf_ = frames[0].sf
                log_sf = float(f32(np.log(f32(sf_[1] / sf_[0]))))
                mark("  [points] C2 views built")
                r2 = self.b.search_local_map_batch(fvs, np.stack([frames[t].Ow for t in c2_t]).astype(f32) if c2_t else np.zeros((0, 3), f32),
                                                   snaps, mof, 0.5, log_sf, 3.0, 0.8)
                for t, (m2, n2, iv) in zip(c2_t, r2):
                    summary[t].update(c2_in_view=iv, c2_matches=n2, c2_sum=self._chk(m2))
                mark("  [points] C2 done")
                return
            fr_ready.wait()
            # ---- C2: ORBmatcher(0.8).SearchByProjection(F, localPoints, th=3) (Tracking.cc:1812) ----
            fvs, mvs, inview = [], [], []
            for t, (mdesc, inv, u, v, xr, lvl, vc) in zip(c2_t, fr):
                if t in cv_of:  # the same frame view as in C3, plus the claims
                    fv = type(cv_of[t]).from_buffer_copy(cv_of[t])
                    keep.append(claimed[t])
                    fv.claimed = claimed[t].__array_interface__["data"][0]
                    fvs.append(fv)
                else:
                    fvs.append(frames[t].view(claimed[t], keep))
                mvs.append(N.make_mappoint_view(mdesc, inv, u, v, xr, lvl, vc, None, keep))
                inview.append(int(inv.sum()))
            mark("  [points] C2 views built")
            r2 = self.b.search_local_points_batch(fvs, mvs, 3.0, 0.8) if batch else [self.b.search_local_points(f, m, 3.0, 0.8) for f, m in zip(fvs, mvs)]
            for t, iv, (m2, n2) in zip(c2_t, inview, r2):
                summary[t].update(c2_in_view=iv, c2_matches=n2, c2_sum=self._chk(m2))
            mark("  [points] C2 done")

        # The point searches and the line side do not depend on each other.  A backend whose line matcher has its own handle
        # (its own stream and staging buffers) lets the point searches run on a second host thread while this one prepares and
        # issues the line searches: the matcher calls block in native code (no GIL), so their device time overlaps the numpy glue
        # of the other side.  Same calls, same inputs, same results — only the order in wall-clock time changes.
        worker = None
        if batch and getattr(self.b, "concurrent_sides", False):
            err = []

            def guarded():
                try:
                    point_searches()
                except BaseException as e:  # re-raised on the caller's thread
                    err.append(e)
            # (projecting the local map while C3 runs was measured and loses: its short device calls then queue behind the
            # search's kernels on a GPU the region grower already fills)
            if not fused:
                frustum_of_local_map()
            fr_ready.set()
            worker = threading.Thread(target=guarded)
            worker.start()
        else:
            if not fused:
                frustum_of_local_map()
            fr_ready.set()
            point_searches()
        # ---- line side of the caller state ----
        prep = FrameLite.prepare_lines_batch(frames, depth)  # what the line side can do before the lines are there
        if hasattr(lines, "result"):
            lines = lines.result()
        mark("lines ready")
        FrameLite.attach_lines_batch(frames, lines, depth, prep)
        seq = getattr(frames[0], "_seq", None) if n else None
        if seq is not None:
            for t, c in enumerate(seq["counts"].tolist()):
                summary[t]["n_kl"] = c
            lmaps, where = LocalMap.line_snapshots_dense(seq, n, self.kf_every)
        else:
            where = None
            lmaps = []
            lm = LocalMap()
            for t in range(n):
                summary[t].update(n_kl=len(lines[t][0]))
                lmaps.append((lm.ls, lm.le, lm.lkl, lm.ldesc))
                if t % self.kf_every == 0:
                    lm.add_keyframe_lines(frames[t])
        mark("lines attached, line maps built")
        # ---- D3: LineMatcher(0.9).SearchByProjection(Cur, Last) (Tracking.cc:1247) ----
        if seq is not None:
            has = seq["counts"] > 0
            d3_t = [t for t in c3_t if has[t - 1] and has[t]]
        else:
            d3_t = [t for t in c3_t if len(frames[t - 1].kls) and len(frames[t].kls)]
        lcv, llv = [], []
        cur_view = {}  # the current-frame side of D3 and D5 is the same view
        if seq is not None and d3_t:
            # the same views as below, for all frames at once: the line features are rows of the extractor's dense output
            tt = np.asarray(d3_t, np.int64)
            lcv = self._dense_lineframe_views(seq, tt, frames[0], keep)
            llv = self._dense_mapline_views(seq, tt - 1, keep)
        else:
            for t in d3_t:
                F, last = frames[t], frames[t - 1]
                s3, e3, okl = last.unproject_lines()
                cur_view[t] = N.make_lineframe_view(F.kls, F.ldesc, None, F.Tcw[:3].reshape(-1), self.K, F.bounds, F.size, keep)
                lcv.append(cur_view[t])
                llv.append(N.make_mapline_view(s3, e3, last.kls, last.ldesc, okl, keep))
        mark("D3 views built")

        def d3_search(lcv=lcv, llv=llv):
            rd3 = self.b.line_search_batch(lcv, llv) if batch else [self.b.line_search_batch([c], [l])[0] for c, l in zip(lcv, llv)]
            mark("  [d3] D3 searched")
            self._line_results(summary, d3_t, rd3, "d3")

        d3_worker = None
        two_handles = batch and getattr(self.b, "concurrent_sides", False)
        if two_handles:  # D3 and D5 are independent: D3 runs on its own thread / handle while D5 is prepared and issued
            import threading
            d3_err = []

            def d3_guarded():
                try:
                    d3_search()
                except BaseException as e:
                    d3_err.append(e)
            d3_worker = threading.Thread(target=d3_guarded)
            d3_worker.start()
        else:
            d3_search()
        # ---- D5: LineMatcher(0.8).SearchByProjection(F, localLines) (Tracking.cc:1863) ----
        if where is not None:
            ok5 = (where["hi"] > where["lo"])[where["snap_of"]] & (seq["counts"] > 0)
            d5_t = [t for t in c2_t if ok5[t]]
        else:
            d5_t = [t for t in c2_t if len(lmaps[t][2]) and len(frames[t].kls)]
        lcv, llv = [], []
        snap_view = {}  # frames between two key frames see the same snapshot of the local line map
        if where is not None and d5_t:
            tt = np.asarray(d5_t, np.int64)
            lcv = self._dense_lineframe_views(seq, tt, frames[0], keep)
            llv = self._dense_snapshot_views(where, tt, keep)
        for t in (d5_t if where is None else ()):
            F = frames[t]
            ls, le, lkl, ldesc = lmaps[t]
            lcv.append(cur_view[t] if t in cur_view else N.make_lineframe_view(F.kls, F.ldesc, None, F.Tcw[:3].reshape(-1), self.K, F.bounds, F.size, keep))
            if id(lkl) not in snap_view:
                snap_view[id(lkl)] = N.make_mapline_view(ls, le, lkl, ldesc, np.ones(len(lkl), np.uint8), keep)
            llv.append(snap_view[id(lkl)])
        mark("D5 views built")
        if two_handles:
            rd5 = self.b.line_search_batch(lcv, llv, True)
        else:
            rd5 = self.b.line_search_batch(lcv, llv) if batch else [self.b.line_search_batch([c], [l])[0] for c, l in zip(lcv, llv)]
        self._line_results(summary, d5_t, rd5, "d5")
        mark("D5 done")
        if d3_worker is not None:
            d3_worker.join()
            if d3_err:
                raise d3_err[0]
        if worker is not None:
            worker.join()
            if err:
                raise err[0]
        mark("end")
        return summary


class GpuBackend:
    """The CUDA library through the C ABI (host buffers)."""

    def __init__(self, api, rows, cols, nfeatures=1000, chunk=32, device=0):
        self.api = api
        self.orb = api.ORBextractor(nfeatures, 1.2, 8, 20, 7, device=device, max_cols=cols, max_rows=rows, max_batch=chunk)
        self.line = api.LineExtractor(device=device, max_cols=cols, max_rows=rows, max_batch=chunk)
        self.m = api.DescriptorMatcher(device=device)
        self.ml = api.DescriptorMatcher(device=device)   # the line side's own handle: stream + staging buffers
        self.ml2 = api.DescriptorMatcher(device=device)  # D3 and D5 are independent calls: a handle each lets them run side by side
        self.concurrent_sides = True

    def scale_factors(self):
        return self.orb.GetScaleFactors()

    def extract_orb(self, frames):
        kps, desc, cnt = self.orb.extract_batch(frames)
        return [(kps[i, :cnt[i]], desc[i, :cnt[i]]) for i in range(len(cnt))]

    def extract_lines(self, frames):
        kls, desc, co, cnt = self.line.extract_batch(frames)
        return [(kls[i, :cnt[i]], desc[i, :cnt[i]], co[i, :cnt[i]]) for i in range(len(cnt))]

    def search_last_frame(self, cv, lv, th):
        return self.m.SearchByProjectionLastFrame(cv, lv, th)

    def search_local_points(self, fv, mv, th, nn):
        return self.m.SearchByProjectionLocalPoints(fv, mv, th, nn)

    def project_lines(self, *a):
        return self.m.project_lines(*a)

    def match_lines(self, *a):
        return self.m.match_lines(*a)

    def search_last_frame_batch(self, cvs, lvs, th):
        return self.m.SearchByProjectionLastFrameBatch(cvs, lvs, th)

    def search_local_points_batch(self, fvs, mvs, th, nn):
        return self.m.SearchByProjectionLocalPointsBatch(fvs, mvs, th, nn)

    def search_local_map_batch(self, fvs, ow, maps, map_of_frame, cos_limit, log_sf, th, nn):
        return self.m.SearchLocalMapBatch(fvs, ow, maps, map_of_frame, cos_limit, log_sf, th, nn)

    def line_search_batch(self, cvs, lvs, second=False):
        return (self.ml2 if second else self.ml).SearchLinesByProjectionBatch(cvs, lvs)

    # F rows (Frame glue): UnprojectStereo / IsInFrustum of many frames in one call
    def unproject_batch(self, off, xy, z, rwc, ow, K):
        return self.m.UnprojectBatch(off, xy, z, rwc, ow, K)

    def is_in_frustum_batch(self, *a):
        # on the line side's handle: the caller's thread projects the local map while the point searches (self.m) run on theirs,
        # and the line searches only start after it
        return self.ml.IsInFrustumBatch(*a)
