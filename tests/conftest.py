import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "volumetric-primitives-net_b200")
for p in (REPO, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(os.path.join(REPO, "tests", "golden", "hotpath_golden.npz")))


@pytest.fixture(scope="session")
def golden2():
    """Round-2 vectors (oracle/make_golden.py::round2): a Chamfer case big enough for the tensor-core filter, and the
    backward of obj_to_view_points / rotate_points_forward_x_axis, all produced by the reference's own code."""
    return dict(np.load(os.path.join(REPO, "tests", "golden", "round2_golden.npz")))


@pytest.fixture(scope="session")
def golden_templates():
    return dict(np.load(os.path.join(REPO, "tests", "golden", "templates.npz")))


@pytest.fixture(scope="session")
def c_oracle():
    """ctypes handle of oracle/_build/libvpn_oracle.so (built on demand with gcc)."""
    import ctypes
    import subprocess
    so = os.path.join(REPO, "oracle", "_build", "libvpn_oracle.so")
    if not os.path.isfile(so):
        subprocess.check_call(["make", "-C", os.path.join(REPO, "oracle")])
    lib = ctypes.CDLL(so)
    lib.vpn_oracle_chamfer_nn.restype = ctypes.c_int

    def nn(p1, p2, threads=0):
        p1 = np.ascontiguousarray(p1, dtype=np.float32)
        p2 = np.ascontiguousarray(p2, dtype=np.float32)
        b, p, _ = p1.shape
        m = p2.shape[1]
        m1 = np.empty((b, p), np.float32); i1 = np.empty((b, p), np.int64)
        m2 = np.empty((b, m), np.float32); i2 = np.empty((b, m), np.int64)
        f = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        rc = lib.vpn_oracle_chamfer_nn(f(p1), f(p2), b, p, m, f(m1), f(i1), f(m2), f(i2), threads)
        assert rc == 0
        return m1, i1, m2, i2

    return nn
