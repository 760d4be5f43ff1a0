"""CPU test: the sinf/cosf used by the steered-BRIEF kernel is bit-identical to the host libm (the reference's
std::cos(float)/std::sin(float), src/ORBextractor.cc:113) on a dense sample of [0, 2*pi].  The exhaustive sweep over all
1,086,918,650 floats of the interval was run once while building the repo (0 mismatches, with and without FMA).
The same program checks the logf restatement (MapPoint::PredictScale) on a strided sweep of all normal positive floats;
its exhaustive sweep (2,130,706,432 floats, 0 mismatches, with and without FMA) was run once as well."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sincos_restatement_matches_libm(tmp_path):
    exe = str(tmp_path / "sincos_check")
    subprocess.check_call(["nvcc", "-O2", "-fmad=false", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "sincos_check.cu")])
    out = subprocess.run([exe, "211"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "mismatches 0" in out.stdout
