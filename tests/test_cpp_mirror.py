"""include/suggest_b200.hpp: compiles and links on any box; runs the reference's end-to-end expectations on a B200."""
import os
import subprocess

import pytest

from conftest import GOLDEN, ROOT
from suggest_b200 import _capi, build

BIN = os.path.join(ROOT, "tests", "cpp", "host_mirror_test")


def compile_mirror():
    build.build()
    src = os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp")
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < max(os.path.getmtime(src), os.path.getmtime(_capi.LIB_PATH),
                                                               os.path.getmtime(os.path.join(ROOT, "include", "suggest_b200.hpp"))):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", BIN,
                               _capi.LIB_PATH, "-Wl,-rpath," + os.path.dirname(_capi.LIB_PATH), "-lpthread"])
    return BIN


def test_cpp_mirror_compiles_and_links():
    assert os.path.exists(compile_mirror())


@pytest.mark.gpu
def test_cpp_mirror_runs_reference_expectations():
    out = subprocess.run([compile_mirror(), os.path.join(GOLDEN, "cars.dict")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "host mirror ok" in out.stdout
