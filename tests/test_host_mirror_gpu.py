"""Builds and runs the C++ host-mirror test (tests/cpp/test_host_mirror.cpp): the reference's
SAD/SATD KATs evaluated through rav1e_b200/host/rav1e_b200.hpp -> C ABI -> CUDA."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_cpp_host_mirror(tmp_path):
    exe = str(tmp_path / "host_mirror")
    pkg = os.path.join(ROOT, "rav1e_b200")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests/cpp/test_host_mirror.cpp"),
                           "-o", exe, "-L" + pkg, "-lb200rdo", "-Wl,-rpath," + pkg])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "host mirror ok" in out.stdout


def test_cpp_host_mirror_compiles():
    """CPU-side: the header and its test compile and link against the shared library."""
    pkg = os.path.join(ROOT, "rav1e_b200")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-fsyntax-only", "-Wall",
                           os.path.join(ROOT, "tests/cpp/test_host_mirror.cpp")])
    assert os.path.exists(os.path.join(pkg, "libb200rdo.so"))
