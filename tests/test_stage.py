"""CPU-only: the host staging pool of the library (csrc/orbx_stage.h -- pageable caller buffers <-> pinned memory, copied by
a few threads plus the caller) as a stand-alone C++ unit test built with g++ (no CUDA needed: the header is host-only)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_stager_native(tmp_path):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    exe = str(tmp_path / "stage_test")
    subprocess.check_call([gxx, "-std=c++17", "-O2", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "native", "stage_test.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "STAGE TEST OK" in out.stdout, out.stdout + out.stderr
