"""CPU test: the sinf/cosf used by the steered-BRIEF kernel is bit-identical to the host libm (the reference's
std::cos(float)/std::sin(float), src/ORBextractor.cc:113) on a dense sample of [0, 2*pi].  The exhaustive sweep over all
1,086,918,650 floats of the interval was run once while building the repo (0 mismatches, with and without FMA).
The same program checks the logf restatement (MapPoint::PredictScale) on a strided sweep of all normal positive floats;
its exhaustive sweep (2,130,706,432 floats, 0 mismatches, with and without FMA) was run once as well."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sincos_restatement_matches_libm(tmp_path):
    exe = str(tmp_path / "sincos_check")
    subprocess.check_call(["nvcc", "-O2", "-fmad=false", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "sincos_check.cu")])
    out = subprocess.run([exe, "211"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "mismatches 0" in out.stdout


def test_double_sincos_restatement_equals_host_libm(tmp_path):
    """csrc/pl_glibc_sincos.cuh (rec.dx / rec.dy of LSD's region2rect, the seeds' initial sums) against std::sin / std::cos of this
    host, compiled without contraction.  The restatement follows the FMA build of glibc's s_sin.c, which is what an x86-64 host
    with FMA3 runs; on another host the reference itself computes other last bits, and the test says so instead of failing."""
    exe = str(tmp_path / "glibc_sincos_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "glibc_sincos_check.cpp"), "-lm"])
    r = subprocess.run([exe, "20000000"], capture_output=True, text=True)
    fma = "fma" in open("/proc/cpuinfo").read() if os.path.exists("/proc/cpuinfo") else False
    if not fma:
        pytest.skip("host without FMA3: its libm runs the non-FMA build of s_sin.c: " + r.stdout.strip().splitlines()[-1])
    assert r.returncode == 0, r.stdout
