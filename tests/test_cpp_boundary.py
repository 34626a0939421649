"""The C++ mirror of the reference interface (include/b2q_executor.hpp) compiles against the C ABI (checked on CPU) and
— on a GPU box — passes the GroupByTest-shaped boundary test (tests/cpp/test_executor_boundary.cpp)."""
import os
import subprocess

import pytest

from heavydb_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _compile(tmp_path):
    lib = build.build()
    exe = tmp_path / "test_executor_boundary"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_executor_boundary.cpp"), "-o", str(exe),
                           lib, f"-Wl,-rpath,{os.path.dirname(lib)}"])
    return exe


def test_cpp_mirror_compiles_and_links(tmp_path):
    assert _compile(tmp_path).exists()


def test_cpp_mirror_surface_links(tmp_path):
    """Estimator unit, getNDVEstimator, deleted-column option, ColumnarResults: compiled and linked, not executed; the
    read-out half (resultSetFromStorage over hand-filled storage, a DECIMAL SUM / AVG among the targets) RUNS, on the CPU."""
    lib = build.build()
    exe = tmp_path / "mirror_surface"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "mirror_surface.cpp"), "-o", str(exe),
                           lib, f"-Wl,-rpath,{os.path.dirname(lib)}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "mirror surface links" in out.stdout


@pytest.mark.gpu
def test_cpp_boundary_like_groupbytest(tmp_path):
    exe = _compile(tmp_path)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "boundary test ok" in out.stdout
