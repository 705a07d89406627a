"""The C++ mirror of the reference's include/amg headers: compiles against libamgb.so
without Eigen (CPU check) and passes the reference-test transcription on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "algebraic-multigrid_b200")
EXE = os.path.join(ROOT, "tests", "cpp", "testlib_gpu")


def build():
    src = os.path.join(ROOT, "tests", "cpp", "testlib_gpu.cpp")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-DAMGB_NO_EIGEN", "-I", os.path.join(ROOT, "include"),
                           src, "-o", EXE, "-L", PKG, "-lamgb", "-Wl,-rpath," + PKG])
    return EXE


def test_headers_compile_and_link_without_eigen():
    assert os.path.exists(build())


EXE_EIGEN = os.path.join(ROOT, "tests", "cpp", "testlib_gpu_eigen")


def build_eigen_branch():
    """The AMGB_HAVE_EIGEN branch of the mirror headers (Eigen::SparseMatrix / Eigen::Map / Eigen::Matrix)
    against tests/cpp/eigen_shim -- the real Eigen is not on this image; the shim has its signatures."""
    src = os.path.join(ROOT, "tests", "cpp", "testlib_gpu.cpp")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-DAMGB_USE_EIGEN", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "tests", "cpp", "eigen_shim"),
                           src, "-o", EXE_EIGEN, "-L", PKG, "-lamgb", "-Wl,-rpath," + PKG])
    return EXE_EIGEN


def test_eigen_branch_of_the_headers_compiles():
    assert os.path.exists(build_eigen_branch())


@pytest.mark.gpu
def test_reference_test_transcription_passes_with_eigen_types_on_gpu():
    out = subprocess.run([build_eigen_branch()], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "AMG converged after 35 iterations." in out.stdout and "0 failed" in out.stdout


@pytest.mark.gpu
def test_reference_test_transcription_passes_on_gpu():
    out = subprocess.run([build()], capture_output=True, text=True, timeout=300)
    print(out.stdout[-2000:])
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "SPGS converged after 900 iterations." in out.stdout
    assert "AMG converged after 35 iterations." in out.stdout
    assert "0 failed" in out.stdout
