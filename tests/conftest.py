import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")
    config.addinivalue_line("markers", "needs_reference: needs /root/reference or the prebuilt oracle/_ref/libref_planner.so")


@pytest.fixture(scope="session")
def host_helpers():
    """g++ build of path_planner_b200/csrc/ppe_math.cuh for host-side pinning of the formulas."""
    import ctypes

    out = os.path.join(ROOT, "tests", "_build", "libhost_helpers.so")
    src = os.path.join(ROOT, "tests", "host_helpers.cpp")
    hdr = os.path.join(ROOT, "path_planner_b200", "csrc", "ppe_math.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call(
            [gxx, "-O2", "-fPIC", "-ffp-contract=off", "-shared", "-x", "c++", "-I" + os.path.dirname(hdr), "-o", out, src, "-lm"]
        )
    return ctypes.CDLL(out)
