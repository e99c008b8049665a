"""Builds and runs the C++ replay of the reference's own hot-path unit tests (tests/cpp/test_wfa2_utils.cpp)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "test_wfa2_utils")


def _build():
    from sawfish_b200 import _lib
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "test_wfa2_utils.cpp")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(_lib.SO_PATH)):
        so_dir = os.path.dirname(_lib.SO_PATH)
        subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-o", EXE, src, "-L" + so_dir, "-lsawfish_cuda",
                        "-Wl,-rpath," + so_dir], check=True)
    return EXE


def test_cpp_host_side_vectors():
    r = subprocess.run([_build(), "host"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_reference_unit_tests_on_gpu():
    r = subprocess.run([_build(), "gpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_k2v2_packed_recurrence_on_host():
    """The s16x2 recurrence of the K2 v2 kernel (v2_recur8: funnel-shifted neighbours, saturating insertion increment, all
    three chunk variants) against the scalar recurrence, on the CPU (the DPX intrinsics are emulated by the header)."""
    exe = os.path.join(ROOT, "build", "test_v2_recur")
    src = os.path.join(ROOT, "tests", "cpp", "test_v2_recur.cu")
    hdr = os.path.join(ROOT, "sawfish_b200", "csrc", "sfc_k_align2.cuh")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["/usr/local/cuda/bin/nvcc", "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe, src],
                       check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr
