"""The C++ host API (include/deeppowers_fhe.hpp) compiles against the C ABI and links to libdpfhe.so."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    import deeppowers_b200
    deeppowers_b200.load_library()
    exe = str(tmp_path / "encrypted_batch")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    lib_dir = os.path.join(ROOT, "deeppowers_b200")
    subprocess.check_call([gxx, "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "encrypted_batch.cpp"), "-L", lib_dir, "-ldpfhe",
                           "-Wl,-rpath," + lib_dir, "-o", exe])
    return exe


def test_cpp_example_builds_and_fails_loudly_without_gpu(tmp_path):
    import torch
    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu-marked test")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1
    assert "no CPU fallback" in r.stderr          # std::runtime_error carrying dpfhe_last_error()


@pytest.mark.gpu
def test_cpp_example_runs_on_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "ct-mults per second" in r.stdout
    assert "identical result" in r.stdout          # MultiEvaluator over every visible GPU reproduces the single-device bits
