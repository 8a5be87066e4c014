"""Runs the C++ re-statement of the reference's gtest cases (tests/cpp/test_host_api.cpp) against the
host mirror of PolynomialOptimization<N>.  The --cpu-only half needs no device."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_host_api.cpp")
BIN = os.path.join(ROOT, "tests", "cpp", "test_host_api")


def build_binary():
    from mav_trajectory_generation_b200 import _build
    _build.build_all()
    deps = [SRC, _build.LIB_HOST, _build.LIB_CUDA]
    if not os.path.exists(BIN) or any(os.path.getmtime(d) > os.path.getmtime(BIN) for d in deps):
        subprocess.check_call([
            os.environ.get("CXX", "g++"), "-O1", "-std=c++17", "-I", os.path.join(_build.HOST, "include"), "-I",
            os.path.join(ROOT, "include"), SRC, "-o", BIN, "-L", _build.PKG, "-lmtg_host", "-lmtg_b200",
            "-Wl,-rpath,$ORIGIN/../../mav_trajectory_generation_b200"])
    return BIN


def test_host_api_cpu_only():
    out = subprocess.run([build_binary(), "--cpu-only"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "0 failures" in out.stdout


@pytest.mark.gpu
def test_host_api_full_on_gpu():
    out = subprocess.run([build_binary()], capture_output=True, text=True, timeout=1200)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "0 failures" in out.stdout


def test_readme_example_compiles():
    """examples/readme_example.cpp is the reference README's program with only the include root changed."""
    from mav_trajectory_generation_b200 import _build
    _build.build_all()
    out = os.path.join(ROOT, "tests", "cpp", "readme_example")
    subprocess.check_call([os.environ.get("CXX", "g++"), "-std=c++17", "-I", os.path.join(_build.HOST, "include"), "-I",
                           os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "readme_example.cpp"), "-o", out,
                           "-L", _build.PKG, "-lmtg_host", "-lmtg_b200",
                           "-Wl,-rpath,$ORIGIN/../../mav_trajectory_generation_b200"])
    assert os.path.exists(out)
