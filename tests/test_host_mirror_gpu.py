"""GPU test: the C++ host-side mirrors (package host/*.h) compiled against the C ABI and run like Frame.cc uses them."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fnv(data: bytes) -> int:
    h = 1469598103934665603
    for b in data:
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def test_cpp_host_mirror(api, oracle, synth, tmp_path):
    pkgdir = os.path.join(ROOT, "orb_slam2_modification_with-point-and-line-feature_b200")
    exe = str(tmp_path / "host_mirror_test")
    subprocess.check_call(["g++", "-std=c++14", "-O1", os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp"), "-o", exe,
                           "-L" + pkgdir, "-lplslam", "-Wl,-rpath," + pkgdir])
    img = synth.frame(1000, 640, 480)
    raw = str(tmp_path / "img.gray")
    img.tofile(raw)
    out = subprocess.run([exe, raw, "640", "480"], capture_output=True, text=True, check=True).stdout.splitlines()
    o = oracle.OrbOracle(1000)
    ok, od = o.extract(img)
    lvl3 = o.level_bordered(3)[19:-19, 19:-19]
    expect = (f"orb n={len(ok)} levels=8 sf={float(o.tables()['scale_factors'][7]):.6f} hash={fnv(ok.tobytes() + od.tobytes())} "
              f"pyr7=179x134 px={int(lvl3[5, 7])}")
    assert out[0] == expect
    kls, ldesc, co = oracle.line_extract(img, 80)
    assert out[1] == f"lines n={len(kls)} coeffs={len(kls)} class0={int(kls['class_id'][0])} deschash={fnv(ldesc.tobytes())}"
    assert out[2] == "empty n=0" and out[3] == "selfdist=0"
