"""The C++ host shim (include/pk_world.hpp) compiles against the C ABI with plain g++ and behaves
like physkit::world for the approach scenario of the reference's tests/co/co_tests.cpp:358-382."""
import os
import subprocess

import pytest

import physkit_b200 as pk
from physkit_b200 import build as pk_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "host_shim_test")


def _compile():
    pk_build.build()
    lib_dir = os.path.dirname(pk.library_path())
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "host_shim_test.cpp"), "-o", EXE,
           "-L", lib_dir, "-lpk_collide", f"-Wl,-rpath,{lib_dir}"]
    subprocess.run(cmd, check=True)


def test_host_shim_compiles_and_refuses_without_gpu():
    import torch

    _compile()
    r = subprocess.run([EXE], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 3, r.stdout + r.stderr  # PK_E_NO_DEVICE surfaced as pk::error, no fallback


@pytest.mark.gpu
def test_host_shim_on_gpu():
    _compile()
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host shim ok" in r.stdout
