"""Host-side unit checks of native building blocks that need no GPU: the radix-16 FFT pass algebra (fft16.cuh is
__host__ __device__) and the worker pool of the staged host copies.  Compiled with nvcc for the host and run here."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "dotsocp_b200", "csrc")


def _build_and_run(tmp_path, name, extra_sources, token):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / name)
    cmd = [nvcc, "-O2", "-std=c++17", "-I", CSRC, os.path.join(ROOT, "tests", "host", name + ".cu"), *extra_sources,
           "-o", exe, "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and token in r.stdout, r.stdout + r.stderr


def test_fft16_pass_algebra_on_host(tmp_path):
    _build_and_run(tmp_path, "fft16_host_test", [], "FFT16_HOST_OK")


def test_worker_pool_on_host(tmp_path):
    _build_and_run(tmp_path, "pool_host_test", [os.path.join(CSRC, "hostcopy.cu")], "POOL_HOST_OK")
