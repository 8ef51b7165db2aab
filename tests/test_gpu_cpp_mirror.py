"""GPU suite: builds tests/cpp/test_reference_kats.cpp against include/zk_b200.hpp (the C++ host-side mirror
of the reference API) and runs it; the CPU part only checks that the mirror compiles and links."""
import os
import subprocess

import pytest

from conftest import ROOT

EXE = os.path.join(ROOT, "build", "test_reference_kats")


def _build():
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    so_dir = os.path.join(ROOT, "zk_b200")
    if not os.path.exists(os.path.join(so_dir, "libzk_b200.so")):
        subprocess.run(["make", "-C", ROOT, "-j8"], check=True, capture_output=True)
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_reference_kats.cpp"),
           "-L", so_dir, "-lzk_b200", f"-Wl,-rpath,{so_dir}", "-o", EXE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_cpp_mirror_compiles_and_links():
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_mirror_reference_kats():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ALL C++ MIRROR TESTS PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
