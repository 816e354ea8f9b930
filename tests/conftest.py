import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _ensure(path, cmd):
    if not os.path.exists(path):
        subprocess.run(cmd, check=True, cwd=ROOT)
    return path


@pytest.fixture(scope="session")
def oracle():
    import streamz_oracle
    return streamz_oracle


@pytest.fixture(scope="session")
def oracle_c():
    """The C restatement (oracle/oracle.c), built on demand; test infrastructure only."""
    so = _ensure(os.path.join(ROOT, "oracle", "_build", "liboracle.so"), ["make", "-C", "oracle"])
    lib = ctypes.CDLL(so)
    lib.so_extract.restype = ctypes.c_size_t
    lib.so_resample.restype = ctypes.c_size_t
    lib.so_train_epoch.restype = ctypes.c_size_t
    lib.so_n_windows.restype = ctypes.c_size_t
    return lib


@pytest.fixture(scope="session")
def fftmath():
    so = os.path.join(ROOT, "tools", "_build", "libfftmath_host.so")
    if not os.path.exists(so):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", os.path.join(ROOT, "tools", "fft_math_host.cpp"),
                        "-o", so], check=True)
    return ctypes.CDLL(so)


@pytest.fixture(scope="session")
def native():
    """ctypes binding of the product library; building it needs nvcc but no GPU."""
    _ensure(os.path.join(ROOT, "streamz_b200", "lib", "libstreamz_b200.so"), ["make", "-C", "streamz_b200/csrc", "-j8"])
    from streamz_b200 import _native
    return _native


@pytest.fixture(scope="session")
def sz(native):
    import streamz_b200
    return streamz_b200


@pytest.fixture(scope="session")
def ctx(sz):
    return sz.Context(0)


@pytest.fixture(scope="session")
def golden():
    return {n[:-4]: np.load(os.path.join(GOLDEN, n)) for n in os.listdir(GOLDEN) if n.endswith(".npy")}


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class CNet(ctypes.Structure):
    """Mirror of `so_net` in oracle/oracle.c."""
    _fields_ = [("n_in", ctypes.c_int), ("h1", ctypes.c_int), ("h2", ctypes.c_int), ("n_out", ctypes.c_int)] + \
               [(n, ctypes.c_void_p) for n in ("w1", "b1", "w2", "b2", "w3", "b3")]

    @classmethod
    def from_arrays(cls, arrs):
        w1, b1, w2, b2, w3, b3 = arrs
        s = cls(w1.shape[0], w1.shape[1], w2.shape[1], w3.shape[1], *[a.ctypes.data for a in arrs])
        s._keep = arrs
        return s
