"""The C++ host class (host/vslam_filter.hpp) compiles against include/ekf_b200.h and links with
libekf_b200.so.  CPU: it must fail loudly without a device (no fallback).  GPU: it runs a step."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(pkg, tmp_path):
    lib = pkg.build()
    exe = str(tmp_path / "host_smoke")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "host_class_smoke.cpp"), lib,
                           f"-Wl,-rpath,{os.path.dirname(lib)}", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"])
    return exe


def test_host_class_links_and_refuses_to_run_without_gpu(pkg, tmp_path):
    import torch
    exe = _build(pkg, tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 3 and "no usable device" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_host_class_runs_a_step(gpu_pkg, tmp_path):
    exe = _build(gpu_pkg, tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "gpu ok" in r.stdout, r.stdout + r.stderr
