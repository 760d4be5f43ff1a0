"""Generates the golden fixtures under tests/golden/ with REAL OpenCV (python cv2 4.13.0) primitives.

Run in the build container (needs cv2):   python tests/golden/make_golden.py

The reference cannot be compiled here (it needs the OpenCV C++ SDK, which the image does not have) and ships no
golden vectors of its own, so the pin is: the reference's *algorithm* (src/ORBextractor.cc,
restated below line by line in Python) executed on top of the *real* OpenCV primitives it calls —
cv2.resize / copyMakeBorder / FastFeatureDetector / GaussianBlur / fastAtan2 / createLineSegmentDetector /
BFMatcher.  The C++ oracle (oracle/) and the CUDA path must reproduce these files.

Fixtures written:
  orb_<name>.npz     keypoints (x,y,size,angle,response,octave), descriptors, per-level candidate counts
  lsd_<name>.npz     cv2 LSD (REFINE_ADV) segments + width/prec/nfa
  knn_<n>.npz        cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2) indices/distances
"""
import ctypes
import importlib
import math
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
synth = importlib.import_module("orb_slam2_modification_with-point-and-line-feature_b200.synth")

cv2.setNumThreads(1)
_libm = ctypes.CDLL("libm.so.6")
_libm.cosf.restype = ctypes.c_float
_libm.cosf.argtypes = [ctypes.c_float]
_libm.sinf.restype = ctypes.c_float
_libm.sinf.argtypes = [ctypes.c_float]

f32 = np.float32
EDGE = 19
HALF = 15


def load_pattern():
    txt = open(os.path.join(HERE, "..", "..", "include", "pl_brief_pattern.inc")).read()
    txt = txt[txt.index("*/") + 2:]
    v = [int(t) for t in txt.replace("\n", "").split(",") if t.strip()]
    assert len(v) == 1024
    return np.array(v, np.int32).reshape(256, 4)


def cv_round(v):  # round-half-even on a float32 value
    return int(np.rint(f32(v)))


class PyOrb:
    """ORBextractor restated on cv2 primitives (reference src/ORBextractor.cc:410-470, 765-853, 1043-1132)."""

    def __init__(self, nfeatures, scale_factor, nlevels, ini_th, min_th):
        self.nfeatures, self.nlevels, self.ini_th, self.min_th = nfeatures, nlevels, ini_th, min_th
        sfd = float(f32(scale_factor))  # double member holding the float argument
        self.sf = [f32(1.0)]
        for i in range(1, nlevels):
            self.sf.append(f32(float(self.sf[i - 1]) * sfd))
        self.invsf = [f32(1.0) / s for s in self.sf]
        factor = f32(1.0 / sfd)
        nd = f32(nfeatures) * (f32(1) - factor) / (f32(1) - f32(math.pow(float(factor), float(nlevels))))
        self.per_level = []
        s = 0
        for _ in range(nlevels - 1):
            self.per_level.append(cv_round(nd))
            s += self.per_level[-1]
            nd = f32(nd * factor)
        self.per_level.append(max(nfeatures - s, 0))
        vmax = int(math.floor(float(f32(HALF) * f32(math.sqrt(2.0)) / f32(2) + f32(1))))
        vmin = int(math.ceil(float(f32(HALF) * f32(math.sqrt(2.0)) / f32(2))))
        um = [0] * (HALF + 1)
        for v in range(vmax + 1):
            um[v] = int(np.rint(math.sqrt(HALF * HALF - v * v)))
        v0 = 0
        for v in range(HALF, vmin - 1, -1):
            while um[v0] == um[v0 + 1]:
                v0 += 1
            um[v] = v0
            v0 += 1
        self.umax = um
        self.pattern = load_pattern()

    def pyramid(self, image):
        pyr = []
        for level in range(self.nlevels):
            scale = self.invsf[level]
            w, h = cv_round(f32(image.shape[1]) * scale), cv_round(f32(image.shape[0]) * scale)
            if level != 0:
                prev = pyr[level - 1][EDGE:-EDGE, EDGE:-EDGE]
                r = cv2.resize(prev, (w, h), interpolation=cv2.INTER_LINEAR)
                temp = cv2.copyMakeBorder(r, EDGE, EDGE, EDGE, EDGE, cv2.BORDER_REFLECT_101 | cv2.BORDER_ISOLATED)
            else:
                temp = cv2.copyMakeBorder(image, EDGE, EDGE, EDGE, EDGE, cv2.BORDER_REFLECT_101)
            pyr.append(temp)
        return pyr

    @staticmethod
    def divide(node):
        (ulx, uly, urx, bry), keys = node["box"], node["keys"]
        halfx = int(math.ceil(float(f32(urx - ulx) / f32(2))))
        halfy = int(math.ceil(float(f32(bry - uly) / f32(2))))
        boxes = [(ulx, uly, ulx + halfx, uly + halfy), (ulx + halfx, uly, urx, uly + halfy),
                 (ulx, uly + halfy, ulx + halfx, bry), (ulx + halfx, uly + halfy, urx, bry)]
        ch = [dict(box=b, keys=[], nomore=False) for b in boxes]
        n1urx, n1bry = f32(ulx + halfx), f32(uly + halfy)
        for kp in keys:
            if kp[0] < n1urx:
                (ch[0] if kp[1] < n1bry else ch[2])["keys"].append(kp)
            elif kp[1] < n1bry:
                ch[1]["keys"].append(kp)
            else:
                ch[3]["keys"].append(kp)
        for c in ch:
            if len(c["keys"]) == 1:
                c["nomore"] = True
        return ch

    def distribute(self, keys, minx, maxx, miny, maxy, N):
        n_ini = int(math.floor(float(f32(maxx - minx) / f32(maxy - miny)) + 0.5))  # C round(): half away from zero
        hx = f32(maxx - minx) / f32(n_ini)
        seq = [0]
        nodes = []  # python list used as the std::list, index 0 = front

        def mk(box):
            d = dict(box=box, keys=[], nomore=False, seq=seq[0])
            seq[0] += 1
            return d

        ini = [mk((int(hx * f32(i)), 0, int(hx * f32(i + 1)), maxy - miny)) for i in range(n_ini)]
        nodes.extend(ini)
        for kp in keys:
            ini[int(kp[0] / hx)]["keys"].append(kp)
        nodes = [n for n in nodes if n["keys"]]
        for n in nodes:
            if len(n["keys"]) == 1:
                n["nomore"] = True
        finish = False
        while not finish:
            prev_size = len(nodes)
            n_to_expand = 0
            vsize = []
            new_front = []  # children in creation order; list front = reversed(new_front)
            keep = []
            for n in nodes:
                if n["nomore"]:
                    keep.append(n)
                    continue
                for c in self.divide(n):
                    if c["keys"]:
                        c["seq"] = seq[0]
                        seq[0] += 1
                        new_front.append(c)
                        if len(c["keys"]) > 1:
                            n_to_expand += 1
                            vsize.append(c)
            nodes = new_front[::-1] + keep
            if len(nodes) >= N or len(nodes) == prev_size:
                finish = True
            elif len(nodes) + n_to_expand * 3 > N:
                while not finish:
                    prev_size = len(nodes)
                    prev = sorted(vsize, key=lambda c: (len(c["keys"]), c["seq"]))
                    vsize = []
                    for c in reversed(prev):
                        for cc in self.divide(c):
                            if cc["keys"]:
                                cc["seq"] = seq[0]
                                seq[0] += 1
                                nodes.insert(0, cc)
                                if len(cc["keys"]) > 1:
                                    vsize.append(cc)
                        del nodes[next(i for i, n in enumerate(nodes) if n is c)]
                        if len(nodes) >= N:
                            break
                    if len(nodes) >= N or len(nodes) == prev_size:
                        finish = True
        out = []
        for n in nodes:
            best = n["keys"][0]
            for kp in n["keys"][1:]:
                if kp[2] > best[2]:
                    best = kp
            out.append(best)
        return out

    def ic_angle(self, plane, x, y):
        m01 = m10 = 0
        for u in range(-HALF, HALF + 1):
            m10 += u * int(plane[y, x + u])
        for v in range(1, HALF + 1):
            d = self.umax[v]
            vs = 0
            for u in range(-d, d + 1):
                p, m = int(plane[y + v, x + u]), int(plane[y - v, x + u])
                vs += p - m
                m10 += u * (p + m)
            m01 += v * vs
        return f32(cv2.fastAtan2(float(f32(m01)), float(f32(m10))))

    def descriptor(self, blurred, x, y, angle_deg):
        factor_pi = f32(math.pi / float(f32(180.0)))
        angle = f32(f32(angle_deg) * factor_pi)
        a, b = f32(_libm.cosf(float(angle))), f32(_libm.sinf(float(angle)))
        desc = np.zeros(32, np.uint8)
        for i in range(32):
            val = 0
            for k in range(8):
                x0, y0, x1, y1 = (f32(t) for t in self.pattern[i * 8 + k])
                t0 = blurred[y + cv_round(f32(x0 * b) + f32(y0 * a)), x + cv_round(f32(x0 * a) - f32(y0 * b))]
                t1 = blurred[y + cv_round(f32(x1 * b) + f32(y1 * a)), x + cv_round(f32(x1 * a) - f32(y1 * b))]
                val |= int(t0 < t1) << k
            desc[i] = val
        return desc

    def __call__(self, image):
        pyr = self.pyramid(image)
        all_kps, cand_counts = [], []
        fast_ini = cv2.FastFeatureDetector_create(self.ini_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        fast_min = cv2.FastFeatureDetector_create(self.min_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        for level in range(self.nlevels):
            plane = pyr[level]
            im = plane[EDGE:-EDGE, EDGE:-EDGE]
            min_b = EDGE - 3
            max_bx, max_by = im.shape[1] - EDGE + 3, im.shape[0] - EDGE + 3
            width, height = f32(max_bx - min_b), f32(max_by - min_b)
            ncols, nrows = int(width / f32(30)), int(height / f32(30))
            wcell, hcell = int(math.ceil(float(width / f32(ncols)))), int(math.ceil(float(height / f32(nrows))))
            cand = []
            for i in range(nrows):
                ini_y = min_b + i * hcell
                max_y = ini_y + hcell + 6
                if ini_y >= max_by - 3:
                    continue
                max_y = min(max_y, max_by)
                for j in range(ncols):
                    ini_x = min_b + j * wcell
                    max_x = ini_x + wcell + 6
                    if ini_x >= max_bx - 6:
                        continue
                    max_x = min(max_x, max_bx)
                    roi = np.ascontiguousarray(im[ini_y:max_y, ini_x:max_x])
                    kc = fast_ini.detect(roi)
                    if not kc:
                        kc = fast_min.detect(roi)
                    for k in kc:
                        cand.append((f32(k.pt[0] + j * wcell), f32(k.pt[1] + i * hcell), f32(k.response)))
            cand_counts.append(len(cand))
            kept = self.distribute(cand, min_b, max_bx, min_b, max_by, self.per_level[level]) if cand else []
            size = f32(int(f32(31) * self.sf[level]))
            kps = []
            for (x, y, r) in kept:
                x, y = f32(x + min_b), f32(y + min_b)
                ang = self.ic_angle(plane, cv_round(x) + EDGE, cv_round(y) + EDGE)
                kps.append([x, y, size, ang, r, level])
            all_kps.append(kps)
        out_k, out_d = [], []
        for level in range(self.nlevels):
            kps = all_kps[level]
            if not kps:
                continue
            im = pyr[level][EDGE:-EDGE, EDGE:-EDGE].copy()
            blurred = cv2.GaussianBlur(im, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
            for kp in kps:
                out_d.append(self.descriptor(blurred, cv_round(kp[0]), cv_round(kp[1]), kp[3]))
                if level != 0:
                    kp[0] = f32(kp[0] * self.sf[level])
                    kp[1] = f32(kp[1] * self.sf[level])
                out_k.append(kp)
        k = np.array(out_k, np.float32).reshape(-1, 6)
        return k, np.array(out_d, np.uint8).reshape(-1, 32), np.array(cand_counts, np.int32)


def make_orb(name, img, nfeatures):
    k, d, cc = PyOrb(nfeatures, 1.2, 8, 20, 7)(img)
    np.savez_compressed(os.path.join(HERE, f"orb_{name}.npz"), keypoints=k, descriptors=d, cand_counts=cc)
    print("orb", name, k.shape, cc.tolist())


def make_lsd(name, img):
    lsd = cv2.createLineSegmentDetector(cv2.LSD_REFINE_ADV)
    lines, width, prec, nfa = lsd.detect(img)
    lines = np.zeros((0, 4), np.float32) if lines is None else lines.reshape(-1, 4)
    np.savez_compressed(os.path.join(HERE, f"lsd_{name}.npz"), lines=lines,
                        width=np.asarray(width, np.float64).ravel(), prec=np.asarray(prec, np.float64).ravel(),
                        nfa=np.asarray(nfa, np.float64).ravel())
    print("lsd", name, lines.shape)


def make_knn(n):
    q, t = synth.descriptor_sets(n)
    m = cv2.BFMatcher(cv2.NORM_HAMMING, False).knnMatch(q, t, 2)
    idx = np.array([[mm[0].trainIdx, mm[1].trainIdx] for mm in m], np.int32)
    dist = np.array([[mm[0].distance, mm[1].distance] for mm in m], np.float32).astype(np.int32)
    np.savez_compressed(os.path.join(HERE, f"knn_{n}.npz"), idx=idx, dist=dist)
    print("knn", n, idx.shape)


if __name__ == "__main__":
    which = sys.argv[1:] or ["orb", "lsd", "knn"]
    if "orb" in which:
        make_orb("cfgA_seed1000", synth.frame(1000, 640, 480), 1000)
        make_orb("small_seed7", synth.frame(7, 320, 240), 500)
        make_orb("kitti_seed3000", synth.frame(3000, 1241, 376), 2000)
        flat = np.full((240, 320), 128, np.uint8)
        flat[100:140, 150:200] = 200  # mostly empty cells: exercises the minTh fallback and tiny candidate sets
        make_orb("sparse", flat, 500)
    if "lsd" in which:
        make_lsd("cfgA_seed1000", synth.frame(1000, 640, 480))
        make_lsd("small_seed7", synth.frame(7, 320, 240))
        make_lsd("kitti_seed3000", synth.frame(3000, 1241, 376))
    if "knn" in which:
        for n in (1024, 2048):
            make_knn(n)
