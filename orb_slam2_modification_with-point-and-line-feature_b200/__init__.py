"""B200-native (sm_100a) point+line SLAM front-end: ORB extraction, LSD+LBD line extraction and the Hamming
descriptor searches of wolfcanli/ORB_SLAM2_Modification_with-point-and-line-feature, behind a C ABI
(include/plslam_c.h -> libplslam.so).  This Python package is only the loader / test-and-bench harness.

Because the directory name contains hyphens, import it with
``importlib.import_module("orb_slam2_modification_with-point-and-line-feature_b200")``.
"""
from . import synth  # noqa: F401  (pure numpy; importable without the native library)

__all__ = ["synth", "load_api"]


def load_api():
    """Imports the ctypes API (fails loudly when libplslam.so is missing — there is no fallback)."""
    from . import api
    return api
