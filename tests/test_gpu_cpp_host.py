"""Builds the C++ host layer (include/streamz_rs.hpp, the stand-in for the Rust shim this image cannot compile) against
libstreamz_b200.so; the gpu-marked test runs it: the reference's unit test and the C1 flow through the C ABI."""
import os
import subprocess

import pytest

from conftest import ROOT

EXE = os.path.join(ROOT, "tests", "cpp", "_build", "host_api_test")


def _build(native):
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    lib_dir = os.path.join(ROOT, "streamz_b200", "lib")
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "host_api_test.cpp"),
                    "-L", lib_dir, "-lstreamz_b200", f"-Wl,-rpath,{lib_dir}", "-o", EXE], check=True)


def test_cpp_host_layer_compiles_and_links(native):
    _build(native)
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_host_layer_runs(native):
    _build(native)
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "host_api_test ok" in out.stdout
