"""CPU tests: the C-ABI shared library builds for sm_100a without a GPU, loads, and exports every entry point that
include/plslam_c.h declares (no compute calls here)."""
import ctypes
import importlib
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "orb_slam2_modification_with-point-and-line-feature_b200"


def declared():
    txt = open(os.path.join(ROOT, "include", "plslam_c.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"PL_API\s+[\w\s\*]+?\b(pl_\w+)\s*\(", txt)))


def test_header_declares_the_path():
    names = declared()
    for must in ("pl_orb_create", "pl_orb_extract", "pl_orb_extract_batch", "pl_orb_extract_batch_dev", "pl_orb_pyramid_read",
                 "pl_line_extract", "pl_line_extract_batch", "pl_hamming_knn2", "pl_hamming_candidates", "pl_orb_search_local_points",
                 "pl_orb_search_last_frame", "pl_line_project", "pl_line_match_pairs", "pl_last_error"):
        assert must in names
    assert len(names) >= 40


def test_library_builds_and_exports_every_declared_symbol():
    build = importlib.import_module(PKG + ".build")
    lib_path = build.build()
    lib = ctypes.CDLL(lib_path)
    missing = [n for n in declared() if not hasattr(lib, n)]
    assert not missing, missing
    lib.pl_build_info.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.pl_build_info()
    # only sm_100a code is embedded (no other arch, no PTX-JIT fallback target)
    out = subprocess.run(["cuobjdump", "--list-elf", lib_path], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_product_does_not_touch_the_oracle():
    """The oracle is test infrastructure: nothing under the package may import / link / reference it."""
    pkgdir = os.path.join(ROOT, PKG)
    for dp, dn, fn in os.walk(pkgdir):
        if "build" in dp.split(os.sep):
            continue
        for f in fn:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "pyoracle" not in src and "libploracle" not in src and "orc_" not in src, os.path.join(dp, f)


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    native = importlib.import_module(PKG + "._native")
    monkeypatch.setattr(native, "_lib", None)
    monkeypatch.setattr(native, "LIB_PATH", str(tmp_path / "nope.so"))
    try:
        native.lib()
        raise AssertionError("expected ImportError")
    except ImportError as e:
        assert "no CPU or PyTorch fallback" in str(e)


def test_cpp_host_mirrors_compile_and_link(tmp_path):
    """Every header-only C++ mirror (ORBextractor, LineExtractor, ORBmatcher / LineMatcher incl. the E rows, ORBVocabulary,
    FrameGlue) compiles as C++14 against the OpenCV-free stand-ins and links against the C-ABI library: each method's call into
    the ABI has the declared signature.  (Nothing is executed: no GPU here.)"""
    build = importlib.import_module(PKG + ".build")
    build.build()
    host = os.path.join(ROOT, PKG, "host")
    src = tmp_path / "mirrors.cpp"
    src.write_text('#include "ORBextractor.h"\n#include "LineExtractor.h"\n#include "Matchers.h"\n#include "FrameGlue.h"\n'
                   "int use(ORB_SLAM2::ORBmatcher& m, ORB_SLAM2::ORBVocabulary& v, ORB_SLAM2::FrameGlue& g) {\n"
                   "  std::vector<int> out; std::vector<std::pair<size_t, size_t>> pairs; std::vector<cv::Point2f> prev;\n"
                   "  pl_frame_view f{}; pl_posepoint_view p{}; pl_triang_view t{}; float z[12] = {0};\n"
                   "  DBoW2::BowVector bv; DBoW2::FeatureVector fv; std::vector<cv::KeyPoint> k, ku; std::vector<float> a, b;\n"
                   "  int n = m.Fuse(f, p, z, 0.18f, z, 3.f, out) + m.Fuse(f, z, 0.18f, p, 3.f, out) + m.SearchBySim3(f, f, p, p, z, z, 0.18f, 0.18f, 7.5f, out);\n"
                   "  n += m.SearchForInitialization(f, f, prev, out, 100) + m.SearchForTriangulation(t, t, z, z, z, 1, 1, 0, 0, z, z, 8, pairs, false);\n"
                   "  m.ComputeDistinctiveDescriptors(nullptr, std::vector<int>{0}, out);\n"
                   "  v.transform(nullptr, 0, bv, fv, 4); g.UndistortKeyPoints(k, 1, 1, 0, 0, z, ku); g.ComputeStereoFromRGBD(z, 1, 1, 4, k, ku, 40.f, a, b);\n"
                   "  return n + (v.empty() ? 1 : 0);\n}\nint main() { return 0; }\n")
    exe = tmp_path / "mirrors"
    r = subprocess.run(["g++", "-std=c++14", "-Wall", "-I", host, "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        "-L" + os.path.join(ROOT, PKG), "-lplslam", "-Wl,-rpath," + os.path.join(ROOT, PKG)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_matcher_shims_with_reference_signatures_compile_and_link(tmp_path):
    """host/shim/ORBmatcher.{h,cc} and LineMatcher.{h,cc} — the classes with exactly the reference's method signatures
    (include/ORBmatcher.h:50-208, include/LineMatcher.h:35-86) — compile against the stand-in Frame / KeyFrame / MapPoint / MapLine
    and link against the C-ABI library together with the driver the GPU test runs.  A call through every reference signature is
    type-checked here (nothing is executed: no GPU)."""
    build = importlib.import_module(PKG + ".build")
    build.build()
    shim = os.path.join(ROOT, PKG, "host", "shim")
    src = tmp_path / "calls.cpp"
    src.write_text('#include "ORBmatcher.h"\n#include "LineMatcher.h"\nusing namespace ORB_SLAM2;\n'
                   "int calls(ORBmatcher& m, LineMatcher& l, Frame& F, Frame& G, KeyFrame* k1, KeyFrame* k2, cv::Mat S) {\n"
                   "  std::vector<MapPoint*> pts, out; std::set<MapPoint*> found; std::vector<MapLine*> lines, lout;\n"
                   "  std::vector<cv::Point2f> prev; std::vector<int> m12; std::vector<std::pair<size_t, size_t>> pairs;\n"
                   "  std::vector<LineMatcher::KeyLine> kls; std::vector<std::pair<int, int>> idx; const float s12 = 1.f;\n"
                   "  int n = m.SearchByProjection(F, pts) + m.SearchByProjection(F, pts, 3.f) + m.SearchByProjection(F, G, 15.f, false);\n"
                   "  n += m.SearchByProjection(F, k1, found, 10.f, 100) + m.SearchByProjection(k1, S, pts, out, 10);\n"
                   "  n += m.SearchByBoW(k1, F, out) + m.SearchByBoW(k1, k2, out) + m.SearchForInitialization(F, G, prev, m12, 100);\n"
                   "  n += m.SearchForTriangulation(k1, k2, S, pairs, false) + m.SearchBySim3(k1, k2, out, s12, S, S, 7.5f);\n"
                   "  n += m.Fuse(k1, pts) + m.Fuse(k1, S, pts, 4.f, out) + ORBmatcher::DescriptorDistance(S, S);\n"
                   "  n += l.SearchByProjection(F, G) + l.SearchByProjection(F, G, kls, idx) + l.SearchByProjection(F, k1, lout) + l.SearchByProjection(F, k1);\n"
                   "  n += l.SearchByProjection(F, k1, kls, idx) + l.SearchByProjection(F, lines) + l.SearchByProjection(F, lines, kls, idx);\n"
                   "  n += l.SearchForTriangulation(k1, k2, pairs, false) + l.Fuse(k1, lines) + LineMatcher::DescriptorDistance(S, S);\n"
                   "  return n + ORBmatcher::TH_LOW + ORBmatcher::TH_HIGH + ORBmatcher::HISTO_LENGTH;\n}\n")
    exe = tmp_path / "shims"
    r = subprocess.run(["g++", "-std=c++14", "-Wall", "-I", shim, str(src), os.path.join(shim, "ORBmatcher.cc"), os.path.join(shim, "LineMatcher.cc"),
                        os.path.join(ROOT, "tests", "cpp", "matcher_shim_test.cpp"), "-o", str(exe),
                        "-L" + os.path.join(ROOT, PKG), "-lplslam", "-Wl,-rpath," + os.path.join(ROOT, PKG)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
