"""Builds and runs the C++ drop-in wrapper test (tests/host_wrapper_test.cpp: ORB_SLAM2::ORBextractor and
ORB_SLAM2::ORBmatcher with the reference's signatures on test doubles of Frame/MapPoint) on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp):
    exe = os.path.join(tmp, "host_wrapper_test")
    subprocess.run(["g++", "-std=c++17", "-O2", "-o", exe, os.path.join(ROOT, "tests", "host_wrapper_test.cpp"),
                    "-L", os.path.join(ROOT, "swarmmap_b200"), "-lswm_orb",
                    "-Wl,-rpath," + os.path.join(ROOT, "swarmmap_b200")], check=True)
    return exe


def test_cpp_wrappers_compile_and_link(swm, tmp_path):
    assert os.path.exists(_build(str(tmp_path)))


@pytest.mark.gpu
def test_cpp_wrappers_run(swm, tmp_path):
    exe = _build(str(tmp_path))
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "HOST_WRAPPER_OK" in r.stdout, r.stdout + r.stderr
