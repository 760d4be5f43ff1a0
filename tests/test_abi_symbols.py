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
