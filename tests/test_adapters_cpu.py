"""The header-only C++ adapters (adapters/*.hpp: InferenceBackend / EgoLanes*Engine lookalikes)
compile against stub OpenCV/reference headers, link to libvp_b200.so, and keep the reference's
error contract (constructors throw std::runtime_error) on a box without a GPU."""
import os
import subprocess

import pytest

from autoware_vision_pilot_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_adapter_check(tmpdir) -> str:
    exe = os.path.join(str(tmpdir), "adapter_check")
    libdir = os.path.dirname(L.LIB_PATH)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", os.path.join(ROOT, "tests", "cpp", "adapter_check.cpp"),
           "-o", exe, "-L" + libdir, "-lvp_b200", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_adapters_compile_link_and_throw(tmp_path):
    L.lib()
    exe = build_adapter_check(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert "ADAPTER_CTOR_THROWS 2" in r.stdout, r.stdout + r.stderr
    assert r.returncode == 0
