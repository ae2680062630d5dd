"""The C++ mirror of the reference API (include/zigz_host.hpp): compiled here, run on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_reference_style.cpp")
BIN = os.path.join(ROOT, "tests", "cpp", "_build", "test_reference_style")


def _build(zlib, po):
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    odir = os.path.join(ROOT, "oracle", "_build")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", SRC, "-o", BIN, "-L" + os.path.dirname(zlib.LIB_PATH), "-lzigz_b200", "-L" + odir,
           "-lzigz_oracle", "-Wl,-rpath," + os.path.dirname(zlib.LIB_PATH), "-Wl,-rpath," + odir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_cpp_api_compiles_and_links(zlib, po):
    _build(zlib, po)
    assert os.path.exists(BIN)


@pytest.mark.gpu
def test_cpp_api_runs_reference_style_tests(zlib, po):
    _build(zlib, po)
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "all checks passed" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
